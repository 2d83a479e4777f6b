"""Bring-up helper (GPU box): run each tcgen05 GEMM operand-major combination in its own process so that a
device trap in one variant does not poison the others.  Prints one line per variant."""
import subprocess
import sys

CODE = r'''
import sys, torch
sys.path.insert(0, ".")
from tests.test_gemm_gpu import _run
a_mn, b_mn, bn, m, n, k = [int(v) for v in sys.argv[1:7]]
err, scale = _run(torch.bfloat16, m, n, k, bool(a_mn), bool(b_mn), bn)
torch.cuda.synchronize()
print(f"RESULT a_mn={a_mn} b_mn={b_mn} bn={bn} m={m} n={n} k={k} err={err:.4g} scale={scale:.4g} ok={err <= 2e-3*scale+1e-3}")
'''

def main():
    cases = []
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            cases.append((a_mn, b_mn, 128, 128, 128, 64))     # one k-block, one tile
            cases.append((a_mn, b_mn, 128, 256, 256, 256))    # several tiles / k-blocks
    cases.append((0, 0, 64, 128, 64, 64))
    cases.append((0, 0, 256, 128, 256, 64))
    for c in cases:
        try:
            r = subprocess.run([sys.executable, "-c", CODE] + [str(v) for v in c], capture_output=True, text=True,
                               timeout=180)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(lines[0] if lines else f"FAIL {c} rc={r.returncode} :: {r.stderr.strip().splitlines()[-1:]}")
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT {c}")
        sys.stdout.flush()

if __name__ == "__main__":
    main()
