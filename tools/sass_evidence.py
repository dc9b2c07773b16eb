"""Build-time evidence (no GPU needed): per kernel of librf_b200.so the register count, spill / stack bytes, static
shared memory (cuobjdump -res-usage) and the counts of the SASS mnemonics that prove the Blackwell path
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, bulk copies -> UBLKCP / UTMALDG, mbarrier ->
SYNCS, legacy tensor path -> HMMA).  Writes profiles/r02s4_sass_evidence.txt (the round-1 state is kept in
profiles/r01_sass_evidence.txt).

    python tools/sass_evidence.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "retrieval-fuse_b200", "librf_b200.so")
OUT = os.path.join(ROOT, "profiles", "r02s4_sass_evidence.txt")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "DFMA", "DADD", "DMUL",
             "LDG", "STG", "LDS", "STS", "SHFL", "FFMA", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
                    break
            counts[cur]["_total"] += 1
    names = demangle(sorted(usage))
    with open(OUT, "w") as f:
        f.write("# cuobjdump -res-usage / -sass of retrieval-fuse_b200/librf_b200.so (sm_100a), per kernel\n")
        f.write("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS\n\n")
        f.write(f"{'kernel':58s} {'regs':>5s} {'stack':>6s} {'smem':>7s} {'instr':>7s}  Blackwell / notable mnemonics\n")
        for mangled in sorted(usage, key=lambda k: short(names[k])):
            u, c = usage[mangled], counts.get(mangled, {})
            notable = " ".join(f"{k}={c[k]}" for k in MNEMONICS[:13] if c.get(k))
            f.write(f"{short(names[mangled])[:58]:58s} {u.get('REG', '?'):>5s} {u.get('STACK', '?'):>6s} {u.get('SHARED', '?'):>7s} "
                    f"{c.get('_total', 0):7d}  {notable}\n")
        tc = [short(names[k]) for k in usage if counts.get(k, {}).get("UTCHMMA")]
        f.write("\n# kernels issuing tcgen05.mma (UTCHMMA): " + ", ".join(sorted(set(tc))) + "\n")
        f.write("# kernels with a non-zero stack frame (local memory): " +
                (", ".join(sorted({short(names[k]) for k, u in usage.items() if int(u.get('STACK', 0)) > 0})) or "none") + "\n")
        legacy = sorted({short(names[k]) for k in usage if counts.get(k, {}).get("HMMA")})
        f.write("# kernels on the legacy mma.sync path (HMMA): " + (", ".join(legacy) or "none") + "\n")
    print(open(OUT).read())


if __name__ == "__main__":
    main()
