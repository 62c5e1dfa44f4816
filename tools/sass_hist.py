"""Static SASS opcode histogram of one kernel of a cubin-carrying binary.

    python tools/sass_hist.py <binary> <regex on the demangled kernel name> [--dump out.sass]

Counts every instruction of the function body (hot and cold paths alike); the executed mix
comes from the ncu captures (profiles/).  Used for before/after comparisons in profiles/."""
import collections
import re
import subprocess
import sys


def functions(binary):
    txt = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
    name, body = None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        yield name, body


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    return out.splitlines()


def main():
    binary, pattern = sys.argv[1], re.compile(sys.argv[2])
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    fns = list(functions(binary))
    pretty = demangle([n for n, _ in fns])
    for (name, body), pn in zip(fns, pretty):
        if not pattern.search(pn):
            continue
        ops = collections.Counter()
        for line in body:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        print("### %s\n%d instructions, %d FP64 (%.0f%%)" % (pn.split("(")[0], total, fp64, 100.0 * fp64 / max(total, 1)))
        print(", ".join("%s %d" % kv for kv in ops.most_common(40)))
        if dump:
            open(dump, "w").write("\n".join(body))


if __name__ == "__main__":
    main()
