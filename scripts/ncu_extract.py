"""Turn `ncu --set full` captures of bench.py into profiles/r2_ncu_metrics.json (+ one CSV per capture under profiles/).

    python scripts/ncu_extract.py gpurun_out/prof_r2_*.ncu-rep

For every kernel the profiled launches are averaged (per-launch values kept alongside): duration, dram__bytes_read/write.sum, smsp__inst_executed.sum, issue
active, achieved occupancy, registers, FMA / LSU pipe utilisation, top stall reasons.  The JSON carries the SHA-256 of the
kernel sources (abacusutils_b200/csrc/abk_*.cu*) it was captured from; bench.py only reports these numbers while the
sources still hash to the same value.  Run it right after the capture, before touching the kernels."""
import csv
import hashlib
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
NAMES = {'tsc_tile_walk_kernel': 'tsc_tile_deposit', 'tsc_bucket_kernel<1': 'tsc_bucket_scatter', 'tsc_bucket_kernel<(bool)1': 'tsc_bucket_scatter', 'tsc_bucket_kernel<0': 'tsc_bucket_hist', 'tsc_bucket_kernel<(bool)0': 'tsc_bucket_hist',
         'power_bin': 'power_bin', 'normalize_kernel': 'normalize_field', 'transpose_scatter_p2p': 'transpose_scatter_p2p'}
KEYS = {'gpu__time_duration.sum': 'duration_ns', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
        'smsp__inst_executed.sum': 'warp_inst_per_launch', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct', 'launch__registers_per_thread': 'registers',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'lsu_pipe_pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_pct',
        'launch__shared_mem_per_block_dynamic': 'smem_dynamic', 'launch__occupancy_limit_shared_mem': 'occ_limit_smem',
        'launch__occupancy_limit_registers': 'occ_limit_regs'}
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e6, 'us': 1e3, 'ns': 1.0, 's': 1e9, 'msecond': 1e6, 'usecond': 1e3, 'nsecond': 1.0, 'second': 1e9}


def num(v):
    try:
        return float(v.replace(',', ''))
    except ValueError:
        return None


def main(reps):
    h = hashlib.sha256()
    for name in ('abk_common.cuh', 'abk_tsc.cu'):      # the sources of the profiled kernels (deposit, bucketing)
        h.update((ROOT / 'abacusutils_b200' / 'csrc' / name).read_bytes())
    out = {'csrc_sha256': h.hexdigest(), 'captures': [Path(r).name for r in reps], 'kernels': {}}
    for rep in reps:
        txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        (ROOT / 'profiles' / (Path(rep).stem + '_raw.csv')).write_text(txt)
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            kn = d.get('Kernel Name', '')
            name = next((v for k, v in NAMES.items() if k in kn), None)
            if name is None:
                continue
            rec = {'kernel': kn[:120], 'grid': d.get('Grid Size'), 'block': d.get('Block Size')}
            for k, short in KEYS.items():
                if k in d and num(d[k]) is not None:
                    rec[short] = num(d[k]) * UNIT.get(units[hdr.index(k)], 1.0)
            rec['dram_bytes_per_launch'] = rec.get('dram_read', 0.0) + rec.get('dram_write', 0.0)
            stalls = sorted(((num(d[k]), k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')) for k in hdr
                             if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio') and num(d[k]) is not None), reverse=True)
            rec['top_stalls_per_issue'] = {k: round(v, 2) for v, k in stalls[:5]}
            out['kernels'].setdefault(name, []).append(rec)
    # one entry per kernel: the mean over the profiled launches (the two tile deposits of a step are different template
    # instances -- plain and widened cell domain -- with different costs; the bench reports their average launch too)
    for name, recs in list(out['kernels'].items()):
        mean = dict(recs[-1])
        for k, v in recs[-1].items():
            if isinstance(v, float):
                mean[k] = sum(r.get(k, v) for r in recs) / len(recs)
        mean['launches_profiled'] = len(recs)
        mean['per_launch'] = [{'kernel': r['kernel'][:70], 'duration_ns': r.get('duration_ns'), 'warp_inst_per_launch': r.get('warp_inst_per_launch'),
                               'dram_bytes_per_launch': r.get('dram_bytes_per_launch'), 'issue_active_pct': r.get('issue_active_pct'),
                               'fma_pipe_pct': r.get('fma_pipe_pct')} for r in recs]
        mean.pop('top_stalls_per_issue', None)
        mean['top_stalls_per_issue'] = recs[-1]['top_stalls_per_issue']
        out['kernels'][name] = mean
    (ROOT / 'profiles' / 'r2_ncu_metrics.json').write_text(json.dumps(out, indent=1))
    for k, v in out['kernels'].items():
        print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ('kernel',)})


if __name__ == '__main__':
    main(sys.argv[1:])
