timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or config0 or config1_surface or edge_cases or exact_cull or boundary_b2 or big_rects or config2 or overflow" 2>&1 | tail -3
bash profiles/run_variants.sh base rows base rows
