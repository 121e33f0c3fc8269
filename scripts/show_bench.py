#!/usr/bin/env python
"""Print the key figures of a bench.py JSON line (default: the last line of gpurun_out/bench.log)."""
import json, sys
f = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.log'
line = [l for l in open(f) if l.startswith('{')][-1]
d = json.loads(line)
print('headline value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 1), d['clocks'])
r = d['roofline']
print('  frac', round(r['frac'], 3), 'share', round(r['kernel_share_of_step'], 4), 'avg launch ms', round(r['avg_launch_ms'], 4), 'redone', r['tiles_re_evaluated_with_full_plan'], '/', r['tiles'],
      'issued frac', round(r['issued']['frac_of_peak'], 3), 'alive', r['issued']['alive_chunks_per_layer_of_8'])
p = d.get('parity')
if p:
    print('  parity ok', p['ok'], 'replay', ['%.1e' % e for e in p['step_replay']['per_step']], 'obj', p['after_30_iterations']['objective_rel_diff'], p['after_200_iterations']['objective_rel_diff'])
    print('  traj 5:', p['after_5_iterations'], '\n  200:', {k: v for k, v in p['after_200_iterations'].items()})
j = d.get('joint')
if j:
    jr = j['roofline']
    print('joint value', round(j['value'], 2), 'e2e', round(j['e2e']['value'], 2), 'ms/step', round(j['ms_per_step'], 1), j['clocks'])
    print('  frac', round(jr['frac'], 3), 'share', round(jr['kernel_share_of_step'], 4), 'fwd ms', round(jr['forward_ms'], 1), 'jac ms', round(jr['jacobian_ms'], 1), 'grad-only ms', round(jr.get('gradient_only_ms', 0), 1),
          'redone', jr['tiles_re_evaluated_with_full_plan'], '/', jr['tiles'], 'issued frac', round(jr['issued']['frac_of_peak'], 3))
    print('  rows/fruit-iter', j['rows_per_fruit_iteration'], 'host==device', j['host_equals_device'], 'status', j['status_bits_seen'])
    jp = j.get('parity')
    if jp:
        print('  parity ok', jp['ok'], 'H', '%.2e' % jp['rel_H'], 'b', '%.2e' % jp['rel_b'], 'flips', jp['membership_flips'], 'rows', jp['rows_forward'], jp['rows_forward_plus_gradient'])
    if 'cpu_baseline' in j:
        print('  cpu', j['cpu_baseline']['value'], 'ratio', round(j['e2e']['value'] / j['cpu_baseline']['value']))
if 'cpu_baseline' in d:
    print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'ratio e2e', round(d['e2e']['value'] / d['cpu_baseline']['value']))
