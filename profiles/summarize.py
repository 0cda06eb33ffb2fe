"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
  python profiles/summarize.py full gpurun_out/prof_sgemm_r1.ncu-rep > profiles/r1_sgemm_full.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = OrderedDict()
    for r in rows:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        name = r['Kernel Name'].split('(')[0]
        ns = float(r['Metric Value'].replace(',', ''))
        if r['Metric Unit'] in ('us', 'usecond'): ns *= 1e3
        if r['Metric Unit'] in ('ms', 'msecond'): ns *= 1e6
        a = agg.setdefault(name, [0, 0.0, r['Grid Size'], r['Block Size']])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f'# {path}: {len(rows)} launches, total device time {total/1e6:.3f} ms (cold-cache, serialised: compare SHARES)')
    print(f'{"kernel":60s} {"launches":>8s} {"total_us":>12s} {"avg_us":>10s} {"share":>7s}  grid / block (last)')
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{name[:60]:60s} {a[0]:8d} {a[1]/1e3:12.1f} {a[1]/1e3/a[0]:10.1f} {100*a[1]/total:6.1f}%  {a[2]} / {a[3]}')


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active', 'sm__inst_executed_pipe_tensor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'smsp__cycles_active.avg', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block', 'smsp__warp_issue_stalled', 'sm__cycles_elapsed.avg', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared']


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(header, r))
        print(f"## {d.get('Kernel Name', '?')[:100]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for h, u in zip(header, units):
            if any(h.startswith(k) for k in KEYS):
                print(f'   {h:80s} {d[h]:>18s} {u}')


def traffic(path):
    """dram bytes (read + write) per launch of the captured kernels, as the JSON bench.py reads for roofline.traffic
    (keys = the names libdrb's own profiler uses)."""
    import json
    import re
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    res = {}
    for r in rows[2:]:
        d = dict(zip(header, r))
        u = dict(zip(header, units))
        name = d.get('Kernel Name', '')
        key = re.search(r'(k_[a-z0-9_]+)', name).group(1)
        if key == 'k_umma_gemm':                     # template argument 2: A read MN-major = dW'^T, K-major = dh
            args = re.search(r'k_umma_gemm<([^>]*)>', name).group(1).replace('(int)', '').replace('(bool)', '').split(',')
            key += '_dw' if args[1].strip() in ('1', 'true') else '_dh'
        key = {'k_gather_chunks': 'k_gather', 'k_scatter_chunks': 'k_scatter'}.get(key, key)
        tot = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            tot += float(d[m].replace(',', '')) * scale[u[m]]
        res[key] = int(tot)
    print(json.dumps({'source': f'{path} (ncu --set full --clock-control none, bench.py --steps 2 --warmup 3, ml-20m shape, '
                                'B=4096)', 'dram_bytes_per_launch': res}, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic}[sys.argv[1]](sys.argv[2])
