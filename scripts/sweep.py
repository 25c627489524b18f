"""run bench.py over a few configurations and print one compact line each:  python scripts/sweep.py "B P cudnn" ..."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cfg in sys.argv[1:]:
    B, P, cb = cfg.split()
    env = dict(os.environ, LAVB_CUDNN_BENCHMARK=cb)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "10", "--warmup", "3", "--batch", B, "--pipelines", P,
                        "--no-cpu-baseline"], capture_output=True, text=True, env=env, timeout=300)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"B={B} P={P} cudnn_bench={cb}: {d['value']:.1f} fps  {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.1f}  "
              f"all-umma {d['roofline']['achieved']:.0f} TF/s  heads {d['roofline_heads_conv']['achieved']:.0f} TF/s  "
              f"pillar {d['roofline_pillar']['achieved']:.0f} GB/s", flush=True)
    except Exception as e:
        print(f"B={B} P={P}: FAILED {e}\n{r.stderr[-800:]}", flush=True)
