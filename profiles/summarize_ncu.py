"""Turn gpurun_out/<tag>_*.ncu-rep into profiles/<tag>_ncu_summary.md + profiles/traffic.json (DRAM bytes per launch)
and profiles/<tag>_launch_shares.md (share of the step per kernel from the launch list).
    python profiles/summarize_ncu.py r01_v4
"""
import csv, glob, io, json, os, subprocess, sys
from collections import defaultdict

tag = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size']
MULT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
lines = [f"# ncu `--set full --clock-control none` summaries, tag {tag}", "",
         "Workload: profiles/prof_step.py = 3 JointTrainer steps on BASELINE configs[2] (1M Gaussians, 1920x1080, SH 3); the 3rd",
         "launch of each kernel is captured.  Times under ncu are cold-cache and serialised (compare shares, not absolutes).", ""]
traffic = {}
for rep in sorted(glob.glob(f'gpurun_out/{tag}_k_*.ncu-rep')):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[-1]
    kname = vals[hdr.index('Kernel Name')].split('(')[0].replace('<unnamed>::', '').replace('void ', '').strip()
    lines += [f"## {kname}", "| metric | value | unit |", "|---|---|---|"]
    d = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f"| {w} | {vals[i]} | {units[i]} |")
            d[w] = (vals[i], units[i])
    try:
        tr = sum(float(d[k][0].replace(',', '')) * MULT[d[k][1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        key = kname.replace('k_', '').split('<')[0]
        key = {'blend_fwd2': 'blend_fwd', 'blend_bwd2': 'blend_bwd', 'ssim_fwd': 'loss_fwd', 'ssim_bwd': 'loss_bwd'}.get(key, key)
        traffic[key] = tr
        lines.append(f"\nDRAM traffic per launch: {tr/1e6:.1f} MB\n")
    except Exception as e:
        lines.append(f"\n(traffic unavailable: {e})\n")
open(f'profiles/{tag}_ncu_summary.md', 'w').write("\n".join(lines))
json.dump(traffic, open('profiles/traffic.json', 'w'), indent=1)
lc = f'gpurun_out/{tag}_launches.csv'
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = rows[0]
    kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rows[1:]:
        try:
            v = float(r[mv].replace(',', ''))
        except ValueError:
            continue
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(r[mu], 1.0)
        name = r[kn].split('(')[0].replace('<unnamed>::', '')[:70]
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    out = [f"# Launch list shares, tag {tag} (ncu gpu__time_duration.sum, 3 steps incl. the first warm-up step)", "",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append(f"| {k} | {cnt[k]} | {v:.1f} | {100*v/s:.1f}% |")
    open(f'profiles/{tag}_launch_shares.md', 'w').write("\n".join(out))
    os.system(f'cp {lc} profiles/{tag}_launches.csv')
print("\n".join(lines[:12]))
