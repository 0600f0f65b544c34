"""A/B timing + parity of attention-kernel variants selected by environment switches (they are read once per process, so
every variant runs in its own interpreter).  Usage on a GPU box:
    python tools/attn_ab.py VRAG_ATTN_DEFER=0 VRAG_ATTN_DEFER=1"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys
sys.path.insert(0, %r)
from verbatim_rag_b200 import _native
ctx = _native.default_context(0)
for name, f in (('fp16 ', ctx.bench_attention), ('split', ctx.bench_attention_split)):
    print(name, 'global %%.4f ms  local %%.4f ms  (1024 x 128 tokens: %%.4f / %%.4f)' %% (
        f(1024, 512, -1, iters=10), f(1024, 512, 64, iters=10), f(1024, 128, -1, iters=10), f(1024, 128, 64, iters=10)), flush=True)
"""


def main():
    for variant in sys.argv[1:] or [""]:
        env = dict(os.environ)
        for kv in filter(None, variant.split(",")):
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=300)
        print("[%s]" % variant, flush=True)
        print(r.stdout.strip() if r.returncode == 0 else r.stderr[-600:], flush=True)
        t = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "--timeout", "300", "-k", "attention",
                            "tests/test_gpu_kernels.py", "tests/test_gpu_precise.py"], env=env, capture_output=True, text=True,
                           timeout=600, cwd=ROOT)
        print("   tests:", t.stdout.strip().splitlines()[-1] if t.stdout.strip() else t.stderr[-300:], flush=True)


if __name__ == "__main__":
    main()
