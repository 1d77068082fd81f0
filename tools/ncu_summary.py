#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU): per launch the duration, DRAM bytes, L2 hit rate, tensor-pipe and
L2 / SM throughput.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_ncu_full_X.json "<source note>" """
import csv
import io
import json
import subprocess
import sys

rep, dst, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
lines = [l for l in raw.splitlines() if l.startswith('"')]
rd = csv.reader(io.StringIO('\n'.join(lines)))
hdr = next(rd)
units = next(rd)


def col(*subs):
    for i, h in enumerate(hdr):                  # exact name first (the report also holds prefixed "Triage" variants)
        if h == subs[0]:
            return i
    for i, h in enumerate(hdr):
        if all(s in h for s in subs):
            return i
    return None


C = {'kernel': col('Kernel Name'), 'grid': col('Grid Size'), 'ns': col('gpu__time_duration.sum'),
     'rd': col('dram__bytes_read.sum'), 'wr': col('dram__bytes_write.sum'),
     'tensor': col('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
     'tensor_el': col('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'), 'memtensor': col('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
     'l2hit': col('lts__t_sector_hit_rate.pct'), 'l2thr': col('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
     'smthr': col('sm__throughput.avg.pct_of_peak_sustained_elapsed'), 'dramthr': col('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
     'regs': col('launch__registers_per_thread'), 'smem_wave': col('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')}


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return None


def scale(i, v):
    u = units[i] if i is not None else ''
    if v is None:
        return None
    return v * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1e3, 'ms': 1e6, 'ns': 1.0, 'usecond': 1e3, 'msecond': 1e6,
                'nsecond': 1.0}.get(u, 1.0)


out = []
for r in rd:
    g = lambda k: (num(r[C[k]]) if C[k] is not None else None)
    name = r[C['kernel']].split('(')[0].replace('(anonymous namespace)::', '')
    out.append({'kernel': name, 'grid': r[C['grid']] if C['grid'] is not None else None,
                'us': round(scale(C['ns'], g('ns')) / 1e3, 1), 'dram_read_MB': round(scale(C['rd'], g('rd')) / 1e6, 1),
                'dram_write_MB': round(scale(C['wr'], g('wr')) / 1e6, 1), 'tensor_pipe_active_pct': g('tensor'), 'tensor_pipe_elapsed_pct': g('tensor_el'), 'mem_tensor_active_pct': g('memtensor'), 'l2_hit_pct': g('l2hit'),
                'l2_throughput_pct': g('l2thr'), 'sm_throughput_pct': g('smthr'), 'dram_throughput_pct': g('dramthr'), 'regs': g('regs')})
res = {'source': note, 'launches': out, 'sum_us': round(sum(o['us'] for o in out), 1),
       'dram_read_MB': round(sum(o['dram_read_MB'] for o in out), 1), 'dram_write_MB': round(sum(o['dram_write_MB'] for o in out), 1)}
json.dump(res, open(dst, 'w'), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != 'launches'}))
for o in out:
    print('%-34s %8.1f us  rd %7.1f MB wr %7.1f MB  tensor %5s%%  L2hit %5s%%  L2thr %5s%%  SMthr %5s%%' % (
        o['kernel'][:34], o['us'], o['dram_read_MB'], o['dram_write_MB'], o['tensor_pipe_active_pct'], o['l2_hit_pct'], o['l2_throughput_pct'], o['sm_throughput_pct']))
