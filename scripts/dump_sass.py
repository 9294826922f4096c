"""Regenerate profiles/sass/: one SASS listing per hand-written kernel + the mnemonic table in README.md.

    python scripts/dump_sass.py        (CPU box: needs only cuobjdump and the built .so)
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "murmura_b200", "ops", "_build", "murmura_b200_ext.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KEY = re.compile(r"^(UTCHMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMACCTL|UBLKCP|SYNCS|UCGABAR_\w+|CGAERRBAR|MAPA|LDGMC\S*|LDG\.\S*|LDG|STG\.\S*|STG|LD\.E\S*|ST\.E\S*|"
                 r"ATOM\S*|RED\S*|MEMBAR\S*|MUFU|FFMA|F2FP|SHFL|LDS|STS|QSPC)")
HEADER = """# SASS listings (`cuobjdump -sass murmura_b200/ops/_build/murmura_b200_ext.so`, sm_100a)

One file per hand-written kernel (regenerate with `python scripts/dump_sass.py`); the table counts the mnemonics that prove the Blackwell / NVLink paths: `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, `UBLKCP` = cp.async.bulk (TMA bulk copy), `SYNCS` = mbarrier, `UTCBAR` = tcgen05.commit, `UCGABAR_*` = cluster barrier, `MAPA` / generic `LD.E`/`ST.E` after it = distributed-shared-memory access, `ATOMS…`/`REDS` on shared::cluster = DSMEM reduction, `LDGMC.E.ADD.F32x4…SYS` = multimem.ld_reduce (NVLS in-switch reduction), `LDG.E.NA.128.CONSTANT` = streaming (peer) loads, `LDG.E.STRONG.SYS` / `STG.E.STRONG.SYS` = acquire / release of the cross-GPU epoch flags.

| kernel | SASS lines | key mnemonics |
|---|---|---|
"""


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    os.makedirs(OUT, exist_ok=True)
    funcs = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []
        elif cur is not None and "/*" in line:
            funcs[cur].append(line)
    rows = {}
    for mangled, lines in funcs.items():
        m = re.search(r"_ZN2mb(\d+)", mangled)
        if not m:
            continue
        n = int(m.group(1)); start = mangled.index(m.group(1)) + len(m.group(1))
        name = mangled[start:start + n]
        tmpl = re.search(r"ILi(\d+)E", mangled[start + n:])
        label = name + (f"<{tmpl.group(1)}>" if tmpl else "")
        cnt = collections.Counter()
        ninstr = 0
        for l in lines:
            mm = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][\w\.]*)", l)
            if not mm:
                continue
            ninstr += 1
            op = mm.group(1)
            if KEY.match(op):
                cnt[op] += 1
        with open(os.path.join(OUT, label.replace("<", "_").replace(">", "") + ".sass"), "w") as f:
            body = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines]           # drop the encoding column
            f.write(f"\t\tFunction : {mangled}\n" + "\n".join(l for l in body if l.strip()) + "\n")
        sig = re.compile(r"^(UTC|LDTM|STTM|UTMA|UBLKCP|SYNCS|UCGABAR|CGAERRBAR|MAPA|LDGMC|QSPC|F2FP)|STRONG\.SYS|\.NA\.")
        first = sorted((kv for kv in cnt.items() if sig.search(kv[0])), key=lambda kv: (-kv[1], kv[0]))
        rest = sorted((kv for kv in cnt.items() if not sig.search(kv[0])), key=lambda kv: (-kv[1], kv[0]))
        top = [f"{k}×{v}" for k, v in (first + rest[:max(0, 16 - len(first))])]
        rows[label] = (ninstr, ", ".join(top))
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write(HEADER)
        for k in sorted(rows):
            f.write(f"| `{k}` | {rows[k][0]} | {rows[k][1]} |\n")
    print(f"{len(rows)} kernels → {OUT}")


if __name__ == "__main__":
    sys.exit(main())
