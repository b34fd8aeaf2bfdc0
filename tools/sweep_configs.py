"""Run BASELINE.json's single-GPU configurations through bench.py and collect the JSON lines
(configs 2, 3, 5; config 4's per-GPU shard = 32 clips x 4 s).  Output: profiles/<tag>_configs.jsonl"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNS = [
    ("config2 flowdec_75m 32x2s NFE6 midpoint", ["--batch", "32", "--seconds", "2", "--N", "3", "--solver", "midpoint"]),
    ("config3 flowdec_25s 64x2s NFE6 midpoint", ["--batch", "64", "--seconds", "2", "--N", "3", "--solver", "midpoint", "--variant", "25s"]),
    ("config4 per-GPU shard 32x4s NFE6 midpoint", ["--batch", "32", "--seconds", "4", "--N", "3", "--solver", "midpoint"]),
    ("config5 NFE1 euler 64x2s", ["--batch", "64", "--seconds", "2", "--N", "1", "--solver", "euler"]),
    ("config5 NFE2 midpoint 64x2s", ["--batch", "64", "--seconds", "2", "--N", "1", "--solver", "midpoint"]),
    ("config5 NFE4 midpoint 64x2s", ["--batch", "64", "--seconds", "2", "--N", "2", "--solver", "midpoint"]),
    ("config5 NFE8 midpoint 64x2s", ["--batch", "64", "--seconds", "2", "--N", "4", "--solver", "midpoint"]),
    ("config5 NFE16 midpoint 64x2s", ["--batch", "64", "--seconds", "2", "--N", "8", "--solver", "midpoint"]),
]


def main(tag):
    out = os.path.join(ROOT, "profiles", f"{tag}_configs.jsonl")
    with open(out, "w") as f:
        for name, extra in RUNS:
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-extras"] + extra
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
            try:
                d = json.loads(line)
                d["name"] = name
                f.write(json.dumps(d) + "\n")
                print(name, "->", round(d["value"], 2), "audio-s/s,", round(d["ms_per_step"], 1), "ms/step, conv frac",
                      round(d["roofline"]["frac"], 3))
            except Exception as e:  # noqa
                print(name, "FAILED", e, r.stderr[-500:])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
