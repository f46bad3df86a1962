#!/usr/bin/env python3
"""One line per bench JSON given on the command line."""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d['roofline']
        pv = (d.get('cpu_baseline') or {}).get('parity_vs_gpu')
        print(f.split('/')[-1], 'N=%d' % d['n_gpus'], 'value %.3fM' % (d['value'] / 1e6), 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3fM' % (d['e2e']['value'] / 1e6),
              'frac', r['frac'], 'ef', d['config'].get('ef'), 'rec', d['config'].get('recall_at_10'), 'clk', (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'),
              'spill/ovf', r.get('visited_spill_queries'), r.get('visited_overflow_queries'), 'cpu %.0f' % ((d.get('cpu_baseline') or {}).get('value') or 0), pv)
    except Exception as e:
        print(f, 'ERR', e)
