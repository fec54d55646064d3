"""Dev helper for the GPU box: run each `-m gpu` test function in its own process with a timeout, so a
trapped kernel (dead CUDA context) or a hang cannot take the other checks down.  Output goes to
gpurun_out/isolated.log; exit code is the number of failed groups.

    python tools/gpu_isolated.py [-k substring] [--per-param] [--timeout 180]
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default="")
    ap.add_argument("--per-param", action="store_true")
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--log", default="isolated.log")
    ap.add_argument("--files", nargs="*", default=["tests/test_gpu_kernels.py", "tests/test_gpu_model.py"])
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    r = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + a.files,
                       cwd=ROOT, capture_output=True, text=True)
    ids = [l.strip() for l in r.stdout.splitlines() if "::" in l]
    if not a.per_param:
        seen, groups = set(), []
        for i in ids:
            gname = i.split("[")[0]
            if gname not in seen:
                seen.add(gname); groups.append(gname)
        ids = groups
    ids = [i for i in ids if a.k in i]
    log = open(os.path.join(ROOT, "gpurun_out", a.log), "w")
    failed = 0
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    for i in ids:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-s", "-x", "--no-header", "-p", "no:cacheprovider", i],
                               cwd=ROOT, capture_output=True, text=True, timeout=a.timeout, env=env)
            rc, out = p.returncode, p.stdout[-6000:] + p.stderr[-3000:]
        except subprocess.TimeoutExpired as e:
            rc, out = -9, "TIMEOUT\n" + ((e.stdout or b"")[-3000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-3000:])
        status = "PASS" if rc == 0 else f"FAIL(rc={rc})"
        failed += rc != 0
        line = f"=== {status} {i} ({time.time() - t0:.1f}s)"
        print(line, flush=True)
        log.write(line + "\n")
        if rc != 0:
            log.write(out + "\n")
            print(out[-2500:], flush=True)
        log.flush()
    print(f"{len(ids) - failed}/{len(ids)} groups passed")
    return failed


if __name__ == "__main__":
    sys.exit(main())
