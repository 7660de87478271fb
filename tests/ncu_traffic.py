"""Writes profiles/traffic.json entries from an `ncu --set full` capture (run here, no GPU needed):

  python tests/ncu_traffic.py gpurun_out/x.ncu-rep KERNEL_REGEX SOURCE.cu[+MORE.cuh] GAMES ACTIONS SIMS [launch_index]

dram_bytes_per_launch = dram__bytes_read.sum + dram__bytes_write.sum of the selected launch.  bench.py's
`roofline.traffic` reads the entry back only while csrc/SOURCE.cu is byte-identical to the captured one
(source_sha16) and the launch shape (games, actions, sims) is the benched one."""
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_scale(unit):
  return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, None)


def main():
  rep, regex, source, games, actions, sims = sys.argv[1:7]
  which = int(sys.argv[7]) if len(sys.argv) > 7 else 0
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  head, units, body = rows[0], rows[1], rows[2:]
  col = {h: i for i, h in enumerate(head)}
  sel = [r for r in body if re.search(regex, r[col["Kernel Name"]])]
  r = sel[which]
  total = 0.0
  for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    total += float(r[col[name]].replace(",", "")) * unit_scale(units[col[name]])
  dur = float(r[col["gpu__time_duration.sum"]].replace(",", "")) if "gpu__time_duration.sum" in col else None
  path = os.path.join(REPO, "profiles", "traffic.json")
  doc = json.load(open(path)) if os.path.exists(path) else {"entries": []}
  h = hashlib.sha256()
  for name in source.split("+"):
    h.update(open(os.path.join(REPO, "model-based-rl_b200", "csrc", name), "rb").read())
  kname = re.search(r"(\w+)\s*(<[^>]*>)?\(", r[col["Kernel Name"]]).group(1)
  entry = {"kernel": kname, "games": int(games), "actions": int(actions),
           "sims": int(sims), "dram_bytes_per_launch": total, "capture": os.path.basename(rep), "source": source,
           "source_sha16": h.hexdigest()[:16],
           "duration_under_ncu": dur, "duration_unit": units[col["gpu__time_duration.sum"]] if dur is not None else None}
  doc["entries"] = [e for e in doc["entries"] if (e["kernel"], e["games"], e["actions"], e["sims"]) !=
                    (entry["kernel"], entry["games"], entry["actions"], entry["sims"])] + [entry]
  json.dump(doc, open(path, "w"), indent=1)
  print(entry)


if __name__ == "__main__":
  main()
