"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total, share, average.
usage: launch_summary.py launches.csv [name_regex_of_our_kernels]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
ours = re.compile(sys.argv[2] if len(sys.argv) > 2 else r'msq|prep_|clean_|features_kernel|crop_|masked_sums|scalars_keypoints|angles_flips|'
                  r'filter_kernel|flips_kernel|inpaint|paste|scale_|kalman|track_angles|tracking_prepare|median_blur|temporal_median')
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
    val = float(r[vi].replace(',', ''))
    val = val / 1e3 if r[ui] in ('ns', 'nsecond') else (val * 1e3 if r[ui] in ('ms', 'msecond') else val)
    tot[name] += val
    cnt[name] += 1
mine = {k: v for k, v in tot.items() if ours.search(k)}
s = sum(mine.values())
print(f'{"kernel":58s} {"launches":>8s} {"total_us":>11s} {"share_of_our_kernels":>21s} {"avg_us":>9s}')
for k, v in sorted(mine.items(), key=lambda kv: -kv[1]):
    print(f'{k[:58]:58s} {cnt[k]:8d} {v:11.1f} {v / s:21.3f} {v / cnt[k]:9.1f}')
other = sum(v for k, v in tot.items() if k not in mine)
print(f'(other kernels in the capture -- torch fills / copies of the bench set-up: {other:.1f} us over {sum(c for k, c in cnt.items() if k not in mine)} launches)')
