"""A/B timing of library variants on ONE box (boxes differ by several us): python tools/ab_bh.py name=path ...
Each variant runs tools/step_time.py in its own process, three rounds, alternating."""
import os
import subprocess
import sys

variants = [a.split("=", 1) for a in sys.argv[1:]]
res = {n: [] for n, _ in variants}
for _ in range(3):
    for n, path in variants:
        env = dict(os.environ, EMBEDDINGNET_B200_LIB=os.path.abspath(path))
        out = subprocess.run([sys.executable, "tools/step_time.py"], env=env, capture_output=True, text=True).stdout.split()
        res[n].append(out)
for n, _ in variants:
    print("%-28s mean us per step: %s   median: %s   loss %s" % (n, " ".join(r[0] for r in res[n]), " ".join(r[1] for r in res[n]), res[n][0][2]))
