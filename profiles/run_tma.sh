timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or config0 or config1_surface or edge_cases or boundary_b2 or precomputed or dropin or trainer_step or golden_raster or tracking or big_rects" 2>&1 | tail -5
bash profiles/run_variants.sh base tma tma_mb5 tma
