"""profiles/sass_r<N>.txt: per-kernel census of the Blackwell-native SASS mnemonics in libdpdist_b200.so (cuobjdump -sass)
and the ptxas register / spill / shared-memory table of the same build (dpdist_b200/build/*.ptxas.txt).
    python tools/sass_census.py > profiles/sass_r3.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dpdist_b200", "libdpdist_b200.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS", "USETMAXREG", "HMMA", "FMNMX3",
         "FMUL2", "FADD2", "MUFU"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def clean(name):
    name = name.replace("(anonymous namespace)::", "").replace("dpd::", "")
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print("# SASS census of %s" % os.path.relpath(LIB, ROOT))
    print("cubin architectures: %s" % ", ".join(arch))
    print("linked libraries (ldd): %s" % ", ".join(sorted(
        l.split()[0] for l in subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout.splitlines()
        if l.strip() and not l.strip().startswith("linux-vdso"))))
    print()
    funcs = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            funcs[cur][op.split(".")[0]] += 1
            if op.startswith("UTCHMMA.2CTA"):
                funcs[cur]["UTCHMMA.2CTA"] += 1
            funcs[cur]["_total"] += 1
    print("| kernel | instructions | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    for f, c in funcs.items():
        name = clean(demangle(f))
        print("| %s | %d | %s |" % (name[-60:], c["_total"], " | ".join(str(c.get(w, 0)) for w in WATCH)))
    print()
    print("# ptxas (registers at launch, spills, barriers) per entry function")
    print("| object | kernel | registers | spill stores / loads (bytes) |")
    print("|---|---|---|---|")
    for p in sorted(glob.glob(os.path.join(ROOT, "dpdist_b200", "build", "*.ptxas.txt"))):
        txt = open(p).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers", txt, re.S):
            name = clean(demangle(m.group(1)))
            print("| %s | %s | %s | %s / %s |" % (os.path.basename(p).replace(".o.ptxas.txt", ""), name[-60:], m.group(5), m.group(3), m.group(4)))


if __name__ == "__main__":
    main()
