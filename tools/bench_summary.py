"""Readable digest of one bench.py JSON line: python tools/bench_summary.py gpurun_out/x_bench.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'on-device', round(d['e2e'].get('result_on_device', 0)), 'clocks', d['clocks'])
r = d['roofline']; print('roofline', round(r['achieved'], 1), 'frac', round(r['frac'], 4), 'burst', round(r['frac_of_burst_peak'], 4), 'launch ms', round(r['avg_launch_ms'], 4), 'share', round(r['share_of_step'], 3))
g = d['gemm_family']; print('family tflops', round(g['tflops'], 1), 'frac', round(g['frac_of_sustained_peak'], 4))
f = d['front_end']; print('front end GB/s', round(f['achieved'], 1), 'frac', round(f['frac'], 4), 'ms', round(f['ms_per_step'], 4))
w = d['whole_step']; print('whole tflops', round(w['tflops'], 1), 'sust', round(w['frac_sustained'], 4), 'burst', round(w['frac_burst'], 4))
for k, v in d['kernels'].items(): print('  ', k, round(v['ms_per_step'], 4), round(v['share'], 3))
cb = d.get('cpu_baseline') or {}
print('cpu', cb.get('value'), cb.get('cores'), 'cores; eager ms', d.get('gpu_eager_baseline', {}).get('fp32_tf32_ms_per_step'), d.get('gpu_eager_baseline', {}).get('bf16_autocast_ms_per_step'))
if 'enc1' in d: print('enc1', round(d['enc1']['ms_per_step'], 3), round(d['enc1']['frac_sustained'], 4))
if 'ragged' in d: print('ragged', round(d['ragged']['ms_padding_encoded'], 3), round(d['ragged']['ms_padding_skipped'], 3))
if 'attention_block' in d: print('attn', {k: round(v, 4) for k, v in d['attention_block'].items() if k != 'what'})
for c in d.get('frontend_sweep', {}).get('cases', []): print('  fe', c['mels'], c['n_fft'], round(c['ms'], 4), round(c['hbm_roofline_frac'], 4), round(c['audio_s_per_s'] / 1e6, 2), 'M audio-s/s')
if 'config4' in d: print('cfg4', round(d['config4']['ms_per_step'], 2), round(d['config4']['tflops'], 1), round(d['config4']['frac_sustained'], 4))
if 'strong_scaling' in d: print('cfg5', round(d['strong_scaling']['ms_per_pass'], 2), round(d['strong_scaling']['audio_s_per_s']))
print('launches', d['gpu_launches'])
