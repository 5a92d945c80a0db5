"""Static SASS summary of the library's kernels: instruction count, opcode histogram, tensor / TMA evidence and the
source lines that own the most instructions.

usage: python tools/sass_summary.py [--nfft 400] [--kernel <substring>] [--lines 25] [--all]

Every translation unit of the library is compiled to a cubin with the product flags (wakeword_trainer_home_b200/build.py)
and disassembled with `nvdisasm -g` (needs -lineinfo, which the product build has).  Without --kernel it prints one
line per kernel (the table committed under profiles/); with --kernel the opcode mix and per-line counts of the kernels
whose demangled name contains the substring.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wakeword_trainer_home_b200 import build as B  # noqa: E402

TENSOR = ("HMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS")


def cubins(nffts):
    d = tempfile.mkdtemp(prefix="wwf_sass_")
    nvcc = "nvcc"
    out = []
    jobs = [("wwfeat", [os.path.join(B.CSRC, "wwfeat.cu")])]
    for n in nffts:
        jobs.append((f"feat_{n}", [f"-DWWF_INST_NFFT={n}", os.path.join(B.CSRC, "wwf_feat_inst.cu")]))
    for name, args in jobs:
        cub = os.path.join(d, name + ".cubin")
        r = subprocess.run([nvcc] + [f for f in B.FLAGS if f not in ("-Xcompiler", "-fPIC")] + ["-cubin", "-o", cub] + args,
                           capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        out.append(cub)
    return out


def demangle(names):
    r = subprocess.run(["cu++filt"] + names, capture_output=True, text=True)
    return r.stdout.splitlines() if r.returncode == 0 else names


def parse(cub):
    """-> {mangled: (Counter opcode, Counter (file, line), n_instr)}"""
    txt = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
    res, fn, cur = {}, None, None
    for l in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1)
            res[fn] = [collections.Counter(), collections.Counter(), 0]
            continue
        if fn is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", l)
        if m:
            res[fn][0][m.group(2)] += 1
            res[fn][1][cur] += 1
            res[fn][2] += 1
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, nargs="*", default=[400])
    ap.add_argument("--all", action="store_true", help="every n_fft family")
    ap.add_argument("--kernel", default=None)
    ap.add_argument("--lines", type=int, default=25)
    a = ap.parse_args()
    nffts = B.N_FFTS if a.all else a.nfft
    for cub in cubins(nffts):
        res = parse(cub)
        names = list(res)
        pretty = dict(zip(names, demangle(names)))
        for fn in names:
            ops, lines, n = res[fn]
            nm = re.sub(r"\(.*", "", pretty[fn].replace("(int)", "")).replace("void ", "")
            if n == 0:
                continue
            if a.kernel is None:
                fp2 = ops["FADD2"] + ops["FFMA2"] + ops["FMUL2"]
                tens = {k: v for k, v in ops.items() if k in TENSOR}
                top = ", ".join(f"{k} {v}" for k, v in ops.most_common(6))
                print(f"{nm:58s} {n:6d} instr | packed fp32x2 {100 * fp2 / n:4.1f}% | tensor/async {tens or '-'} | {top}")
            elif a.kernel in nm:
                print(f"== {nm}: {n} SASS instructions")
                print("   " + ", ".join(f"{k} {100 * v / n:.1f}%" for k, v in ops.most_common(24)))
                for (f, ln), c in lines.most_common(a.lines):
                    path = os.path.join(B.CSRC, f or "")
                    src = ""
                    if f and os.path.exists(path):
                        L = open(path).read().splitlines()
                        src = L[ln - 1].strip() if ln <= len(L) else ""
                    print(f"   {c:5d} {100 * c / n:4.1f}%  {f}:{ln}  {src[:110]}")


if __name__ == "__main__":
    main()
