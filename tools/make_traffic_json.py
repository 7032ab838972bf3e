"""profiles/raycast_traffic.json from an ncu capture of the shipped ray-cast kernel (read here, no GPU): DRAM bytes and warp
instructions per launch + the hash of the ray-cast sources the capture belongs to (bench.py quotes it only for that build).
usage: python tools/make_traffic_json.py gpurun_out/prof_X.ncu-rep profiles/rN_ncu_....txt"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isaac_rover_b200 import _build      # noqa: E402

rep, src_note = sys.argv[1], sys.argv[2]
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
h, r = raw[0], raw[2]
g = lambda k: float(r[h.index(k)].replace(",", ""))      # noqa: E731
units = raw[1]
rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rd *= scale[units[h.index("dram__bytes_read.sum")]]
wr *= scale[units[h.index("dram__bytes_write.sum")]]
grid = int(float(r[h.index("launch__grid_size")]))
out = {"kernel": r[h.index("Kernel Name")], "source": "%s (ncu --set full --clock-control none, 4096 envs, grid %d)" % (src_note, grid),
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
       "warp_instructions_per_launch": g("smsp__inst_executed.sum"), "envs_per_launch": 4096,
       "gpu_time_us_under_ncu": g("gpu__time_duration.sum"), "source_hash": _build.raycast_hash()}
json.dump(out, open(os.path.join(ROOT, "profiles", "raycast_traffic.json"), "w"), indent=1)
print(out)
