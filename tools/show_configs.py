"""Dev tool: print gpurun_out/configs.jsonl (tools/bench_configs.py output) as a compact table."""
import json, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/configs.jsonl'
for l in open(path):
    d = json.loads(l)
    print(d['case'], d['tokens'], d['heads'])
    for k, v in d.items():
        if isinstance(v, dict):
            if 'error' in v:
                print('   ', k, 'ERR', v['error'][:100]); continue
            print('    %-10s ms=%.4f GB/s=%.0f hbm=%.2f TF=%.0f tc=%.3f' % (k, v['ms'], v['GBs'], v['hbm_frac'], v['TFLOPs'], v['tc_frac']))
