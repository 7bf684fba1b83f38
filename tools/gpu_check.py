"""Report-everything kernel check for the GPU box (debug aid; the asserting version is tests/test_kernels_gpu.py).
Each group runs in its own subprocess under a timeout so one hung kernel cannot take the whole call down."""
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_inproc(names):
    import torch
    from tests import kernel_checks as KC
    res = {}
    for name, fn, kw in KC.ALL:
        if names and name not in names:
            continue
        t0 = time.time()
        try:
            err, tol, info = fn(**kw)
            torch.cuda.synchronize()
            res[name] = dict(err=err, tol=tol, ok=bool(err <= tol), info=info, s=round(time.time() - t0, 2))
        except Exception as e:  # noqa
            res[name] = dict(err=None, ok=False, exc=repr(e)[:300], tb=traceback.format_exc()[-600:])
        print(name, json.dumps(res[name]), flush=True)
    return res


def main():
    if "--only" in sys.argv:
        names = sys.argv[sys.argv.index("--only") + 1].split(",")
        res = run_inproc(names)
        print("@@RESULT@@" + json.dumps(res))
        return
    from tests import kernel_checks as KC
    names = [n for n, _, _ in KC.ALL]
    gemm_like = lambda n: any(n.startswith(x) for x in ("linear", "conv", "rope", "idx", "final"))  # noqa: E731
    groups = [[n for n in names if not gemm_like(n) and not n.startswith("attn")]]
    groups += [[n] for n in names if gemm_like(n) or n.startswith("attn")]
    if os.environ.get("F3R_CHECK_FILTER"):
        groups = [[n for n in g if os.environ["F3R_CHECK_FILTER"] in n] for g in groups]
        groups = [g for g in groups if g]
    allres = {}
    for g in groups:
        try:
            p = subprocess.run([sys.executable, __file__, "--only", ",".join(g)], capture_output=True, text=True,
                               timeout=int(os.environ.get("F3R_CHECK_TIMEOUT", "75")))
            out = p.stdout
            got = None
            for line in out.splitlines():
                if line.startswith("@@RESULT@@"):
                    got = json.loads(line[len("@@RESULT@@"):])
            if got is None:
                got = {n: dict(ok=False, exc="no result; rc=%d" % p.returncode, tail=(out + p.stderr)[-800:]) for n in g}
            allres.update(got)
        except subprocess.TimeoutExpired as e:
            so = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            for n in g:
                allres[n] = dict(ok=False, exc="TIMEOUT (hang)", tail=so[-500:])
        for n in g:
            print(("OK   " if allres[n].get("ok") else "FAIL ") + n, {k: v for k, v in allres[n].items() if k != "tb"},
                  flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_checks.json"), "w") as f:
        json.dump(allres, f, indent=1)
    bad = [n for n, r in allres.items() if not r.get("ok")]
    print("FAILED:", bad)


if __name__ == "__main__":
    main()
