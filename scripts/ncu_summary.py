"""Compact per-kernel summary of an Nsight Compute report (run where ncu is installed; no GPU needed):

    python scripts/ncu_summary.py gpurun_out/<dir>/kernels.ncu-rep > profiles/<name>.txt
    python scripts/ncu_summary.py --launches gpurun_out/<dir>/launches.csv > profiles/<name>.txt

Per launch: device time, DRAM bytes read / written and the achieved DRAM GB/s, DRAM and tensor-pipe utilisation, resident
warps, registers. DRAM bytes are what reached HBM during the launch: a kernel whose output still sits in the 126 MB L2
shows few written bytes."""
import collections
import csv
import re
import subprocess
import sys


def raw_rows(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    return float(val.replace(',', '') or 0) * mult


def to_us(val, unit):
    mult = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'second': 1e6}.get(unit, 1)
    return float(val.replace(',', '') or 0) * mult


def full(path):
    hdr, units, rows = raw_rows(path)
    idx = {h: i for i, h in enumerate(hdr)}
    tensor_key = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
    print(f"# {path}: ncu --set full --clock-control none (cold caches, serialised launches)")
    print(f"{'kernel':58s} {'grid':>14s} {'block':>11s} {'regs':>4s} {'us':>7s} {'rd MB':>7s} {'wr MB':>7s} {'GB/s':>7s} "
          f"{'dram%':>6s} {'tens%':>6s} {'warps%':>6s}")
    for r in rows:
        g = lambda k: r[idx[k]] if k in idx else ''   # noqa: E731
        name = re.sub(r'\(.*', '', g('Kernel Name')).replace('void ', '').replace('pe::', '').replace('<unnamed>::', '')
        us = to_us(g('gpu__time_duration.sum'), units[idx['gpu__time_duration.sum']])
        rd = to_bytes(g('dram__bytes_read.sum'), units[idx['dram__bytes_read.sum']])
        wr = to_bytes(g('dram__bytes_write.sum'), units[idx['dram__bytes_write.sum']])
        print(f"{name[:58]:58s} {g('Grid Size'):>14s} {g('Block Size'):>11s} {g('launch__registers_per_thread'):>4s} "
              f"{us:7.2f} {rd / 1e6:7.2f} {wr / 1e6:7.2f} {(rd + wr) / us / 1e3:7.0f} "
              f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or 0):6.1f} "
              f"{float(g(tensor_key) or 0):6.1f} {float(g('sm__warps_active.avg.pct_of_peak_sustained_active') or 0):6.1f}")


def launches(path):
    lines = [ln for ln in open(path, encoding='utf-8') if ln.startswith('"')]
    agg, total = collections.OrderedDict(), 0.0
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')
        us = to_us(r['Metric Value'], r['Metric Unit'])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are serialised and "
          "cold-cache - compare SHARES")
    print(f"total {total:.1f} us over {sum(a[0] for a in agg.values())} launches")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / total:6.1f}%  n={n:4d} avg {t / n:8.2f} us  {name[:100]}")


if __name__ == '__main__':
    if sys.argv[1] == '--launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[1])
