"""Executed instructions of one kernel by source function / source line.

    python tools/ncu_lines.py <binary> <kernel regex> <ncu --page source --csv file> [units] [--lines]

ncu's SASS page gives "Instructions Executed" per SASS instruction; `nvdisasm -gi` of the same
binary gives the (inlined) source position of every instruction in the same order.  The two are
joined by instruction index and summed per innermost source function (found by scanning the
source file for the definition that precedes the line), FP64-pipe instructions counted apart.
`units` (evaluations of the captured launch) turns the sums into instructions per 32 units.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
_defs = {}


def func_of(path, line):
    if path not in _defs:
        d = []
        try:
            for i, text in enumerate(open(path, errors="replace"), 1):
                m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:MBB_HD|__device__|__global__|static|inline|__forceinline__|"
                             r"__host__|constexpr|\s)+[\w:<>\*&\s]+?\b(\w+)\s*\([^;]*$", text)
                if m and not text.lstrip().startswith(("//", "return", "if", "for", "while")):
                    d.append((i, m.group(1)))
        except OSError:
            pass
        _defs[path] = d
    name = "?"
    for i, n in _defs[path]:
        if i > line:
            break
        name = n
    return name


def disasm(binary, pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, group, fresh = None, [], True
        for line in txt.splitlines():
            m = re.match(r"\.text\.(\S+):", line)
            if m:
                cur, group, fresh = m.group(1), [], True
                continue
            if line.lstrip().startswith(".section"):
                cur = None
            if cur is None:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                # a group of annotations (innermost position first, then its inline callers)
                # precedes the instructions it applies to
                if fresh:
                    group, fresh = [], False
                group.append((os.path.normpath(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
            if m:
                out.append((cur, m.group(2), group[0] if group else ("?", 0)))
                fresh = True
    names = subprocess.run(["c++filt"], input="\n".join(n for n, _, _ in out), capture_output=True, text=True).stdout.splitlines()
    return [(op, where) for (n, op, where), pn in zip(out, names) if pattern.search(pn)]


def main():
    binary, pattern, path = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
    rest = [a for a in sys.argv[4:] if not a.startswith("--")]
    units = float(rest[0]) if rest else None
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = rows[hdr]
    ie = cols.index("Instructions Executed")
    ncu = [(r[1].split()[0] if not r[1].lstrip().startswith("@") else r[1].split()[1], int(r[ie])) for r in rows[hdr + 1:]]
    dis = disasm(binary, pattern)
    if len(dis) != len(ncu):
        print("instruction count mismatch: nvdisasm %d, ncu %d" % (len(dis), len(ncu)), file=sys.stderr)
    per = collections.defaultdict(lambda: [0, 0])
    lines = collections.defaultdict(lambda: [0, 0])
    for (op, (f, ln)), (op2, n) in zip(dis, ncu):
        key = "%s:%s" % (os.path.basename(f), func_of(f, ln))
        is64 = op.split(".")[0] in FP64
        per[key][0] += n
        per[key][1] += n if is64 else 0
        lk = "%s:%d" % (os.path.basename(f), ln)
        lines[lk][0] += n
        lines[lk][1] += n if is64 else 0
    tot = sum(v[0] for v in per.values())
    t64 = sum(v[1] for v in per.values())
    scale = 32.0 / units if units else 1.0
    what = "per 32 units" if units else "warp instructions"
    print("executed: %.0f (%s), FP64 %.0f (%.1f%%)" % (tot * scale, what, t64 * scale, 100.0 * t64 / max(tot, 1)))
    print("| source function | executed | FP64 | other | share |\n|---|---:|---:|---:|---:|")
    for k, (n, n64) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        if n * 200 < tot:
            continue
        print("| %s | %.1f | %.1f | %.1f | %.1f%% |" % (k, n * scale, n64 * scale, (n - n64) * scale, 100.0 * n / tot))
    if "--lines" in sys.argv:
        print("\n| source line | executed | FP64 |\n|---|---:|---:|")
        for k, (n, n64) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("NCU_LINES_TOP","60"))]:
            print("| %s | %.1f | %.1f |" % (k, n * scale, n64 * scale))


if __name__ == "__main__":
    main()
