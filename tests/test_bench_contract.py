"""bench.py's reference arm (the CPU port of the path on the host cores, the one other place besides tests / smoke that may
execute oracle/) prints the contract's JSON line without touching a GPU; the B200 arm refuses to run without one."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
  return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + list(flags), capture_output=True, text=True,
                        timeout=600, cwd=REPO)


def test_reference_arm_prints_the_contract_line():
  r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--games", "64", "--sims", "5", "--ref-moves-per-step", "1")
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.strip()]
  assert len(lines) == 1  # ONE JSON line on stdout
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "mcts_node_expansions_per_sec" and d["unit"] == "expansions/s"
  assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
  assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
  assert d["config"]["workload"].startswith("C4") and d["config"]["games_per_gpu"] == 64
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
  assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_needs_a_gpu():
  import torch
  if torch.cuda.is_available():
    return
  r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
  assert r.returncode != 0 and "GPU" in (r.stderr + r.stdout)  # no CPU fallback for the product path
