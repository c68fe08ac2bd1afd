"""SASS evidence for profiles/: mnemonic counts per hot kernel of the built library (cuobjdump -sass).

    python tools/sass_evidence.py <tag>     ->  profiles/sass_<tag>.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
so = os.path.join(ROOT, "rl_ptg_b200", "csrc", "libptg_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("(int)", "").replace("(bool)", "").replace("void ", "").split("(")[0]  # noqa: E731
PATTERNS = collections.OrderedDict([
    ("UBLKCP (TMA bulk store)", r"\bUBLKCP"), ("LDG.*256", r"\bLDG\.[A-Z0-9.]*256"), ("LDG ELL2 (evict_last)", r"\bLDG\.[A-Z0-9.]*ELL2"),
    ("LDG.EF (streamed)", r"\bLDG\.E\.EF"), ("STG.EF (streamed)", r"\bSTG\.[A-Z0-9.]*EF"), ("SHFL", r"\bSHFL"), ("ELECT", r"\bELECT"),
    ("FENCE.VIEW.ASYNC", r"FENCE\.VIEW\.ASYNC"), ("UTMACMDFLUSH", r"UTMACMDFLUSH"), ("ACQBULK / DEPBAR bulk", r"ACQBULK|SYNCS"),
    ("HMMA / UTCMMA (tensor)", r"\bHMMA|UTC[A-Z]*MMA"), ("DFMA / DMUL / DADD", r"\bD(FMA|MUL|ADD)"), ("STL / LDL (spills)", r"\b(STL|LDL)")])
# k_step<NV, MOD, MANY, EVAL, PAC, FLAT>: key-major single step / roll-out, flat single step / roll-out
HOT = ["k_step<4, 1, 0, 0, 13, 0>", "k_step<4, 1, 1, 0, 13, 0>", "k_step<4, 1, 0, 0, 13, 1>", "k_step<4, 1, 1, 0, 13, 1>", "k_reset<4, 1, 0>", "k_episode_stats", "k_features", "k_gae", "k_vecnorm_returns"]
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
        funcs[cur].append(line)
md = [f"# SASS evidence ({tag}): `cuobjdump -sass rl_ptg_b200/csrc/libptg_b200.so` (one sm_100a cubin)", "",
      "Instruction counts per hot kernel (static).  `UBLKCP` = `cp.async.bulk.global.shared::cta` (TMA bulk store of a warp's staged",
      "window tile / feature rows), `LDG.*.256` = 256-bit gathers (sm_100+), `ELL2` = L2 `evict_last` as an instruction modifier (tables,",
      "plant state, RNG records), `.EF` = evict-first streaming accesses (actions, observations, rewards), `SHFL` = lane-pair exchanges of",
      "the paired gathers.  No tensor-core instructions: nothing on this path is a contraction.", "",
      "| kernel | instructions | " + " | ".join(PATTERNS) + " |", "|---|---:|" + "---:|" * len(PATTERNS)]
for name, body in funcs.items():
    if not any(name.startswith(h) for h in HOT):
        continue
    text = "\n".join(body)
    md.append(f"| `{name}` | {len(body)} | " + " | ".join(str(len(re.findall(p, text))) for p in PATTERNS.values()) + " |")
# excerpt: the paired hour-row gather of the key-major single-step kernel (two LDG.256 behind one SHFL of the row index)
body = funcs.get("k_step<4, 1, 0, 0, 13, 0>", [])
idx = [i for i, l in enumerate(body) if re.search(r"LDG\.[A-Z0-9.]*ELL2\.256", l)]
md += ["", "## Excerpt: `k_step<4, 1, 0, 0, 13, 0>` (key-major single step), every 256-bit gather, bulk store and fence", "", "```"]
for i, l in enumerate(body):
    if re.search(r"LDG\.[A-Z0-9.]*256|UBLKCP|FENCE\.VIEW|UTMACMDFLUSH|ELECT|STG\.[A-Z0-9.]*EF|LDG\.E\.EF", l):
        md.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
md += ["```"]
out = os.path.join(ROOT, "profiles", f"sass_{tag}.md")
open(out, "w").write("\n".join(md) + "\n")
print(out, len(md), "lines")
