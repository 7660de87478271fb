"""The device computes Node.expand's math.exp (mcts.py:52) with csrc/mz_exp_algo.h.  The same header
compiled as plain C must agree with the host libm bit for bit on float32-valued arguments (strided
sample here; the full 2^32 sweep is `tests/exp_exhaustive 1`, recorded in DESIGN.md)."""
import os
import subprocess
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exp_algorithm_matches_libm_bit_for_bit():
  src = os.path.join(REPO, "tests", "exp_exhaustive.c")
  inc = os.path.join(REPO, "model-based-rl_b200", "csrc")
  with tempfile.TemporaryDirectory() as d:
    exe = os.path.join(d, "exp_exhaustive")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-I", inc, "-o", exe, src, "-lm"])
    bad, n = [int(v) for v in subprocess.check_output([exe, "509"]).split()]
  assert n > 4_000_000
  assert bad == 0, "%d of %d float32 inputs differ from libm exp" % (bad, n)
