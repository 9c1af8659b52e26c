"""SASS evidence for the hot kernels of libcu2b.so (needs no GPU): per kernel the counts of the mnemonics that matter
for the design claims in DESIGN.md plus one excerpt line for each key instruction.
    python tools/sass_excerpts.py > profiles/r2_sass_excerpts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cu2rec_b200", "lib", "libcu2b.so")

# (title, regex on the mangled name)
KERNELS = [
    ("mf_sgd_user_rounds<32,1,0,5> (single-GPU default: fused sampler, P row in registers for a round)", r"mf_sgd_user_roundsILi32ELi1ELi0ELi5E"),
    ("mf_sgd_user_runs<32,1,THIN,LINKED> (DSGD sub-epoch: wait + updates + hand-off in one launch)", r"mf_sgd_user_runsILi32ELi1ELb1ELb1E"),
    ("mf_sgd_hogwild<32,1,...> (iteration-synchronous order, TMA-fed triplet stream)", r"mf_sgd_hogwildILi32ELi1ELi1ELi3E"),
    ("mf_loss_fused<8,4,1> (single-pass RMSE/MAE, k = 128)", r"mf_loss_fusedILi8ELi4ELi1E"),
    ("dsgd_sample_runs_kernel (per-user sampler of a DSGD round)", r"dsgd_sample_runs_kernel"),
    ("predict_candidates_kernel<16,false> (tcgen05, user tile resident, k <= 128)", r"predict_candidates_kernelILi16ELb0E"),
    ("predict_candidates_kernel<16,true> (tcgen05, streamed chunks, k > 128)", r"predict_candidates_kernelILi16ELb1E"),
    ("mf_sgd_blocked_round<32,1> (deterministic conflict-free mode)", r"mf_sgd_blocked_roundILi32ELi1E"),
]
KEEP = re.compile(r"^(REDG|ATOMG|ATOMS|LDG|STG|LDS|STS|LDTM|UTCHMMA|UTCBAR|UTMALDG|UBLKCP|SYNCS|SHFL|FFMA|FADD|FMUL|FMNMX|DFMA|DADD|"
                  r"MEMBAR|ERRBAR|CCTL|NANOSLEEP|BAR|VOTE|REDUX|CREDUX|IMAD\.WIDE|LOP3|LDL|STL)")
EXCERPT = re.compile(r"REDG\.E\.ADD\.F32x4|UTCHMMA|LDTM|UTMALDG|UBLKCP|SYNCS\.ARRIVE\.TRANS64 |SYNCS\.PHASECHK|STRONG\.SYS|MEMBAR\.(SC|ALL)\.SYS|NANOSLEEP|"
                     r"LDG\.E\.128\.STRONG\.GPU|DFMA|CREDUX")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
            funcs[name].append(line.strip())
    print("# SASS evidence for the hot kernels of libcu2b.so (round 2 final build)")
    print("# produced on the build box (no GPU needed) by tools/sass_excerpts.py: cuobjdump -sass cu2rec_b200/lib/libcu2b.so ;")
    print("# per kernel: counts of the memory / tensor / synchronisation mnemonics + one excerpt line per key instruction")
    print("# REDG.E.ADD.F32x4 = red.global.add.v4.f32 (128-bit L2 atomic add of an SGD step); UBLKCP = cp.async.bulk (TMA 1-D);")
    print("# UTMALDG = cp.async.bulk.tensor (TMA 2-D); UTCHMMA = tcgen05.mma kind::tf32; LDTM = tcgen05.ld; SYNCS = mbarrier ops;")
    print("# LDG/STG ... STRONG.SYS + MEMBAR.SC.SYS = the system-scope flag protocol of the NVLink hand-off")
    for title, pat in KERNELS:
        hit = [n for n in funcs if re.search(pat, n)]
        print("\n## " + title)
        if not hit:
            print("   (no such kernel in this build)")
            continue
        n = hit[0]
        ins = funcs[n]
        print("   symbol " + n)
        print("   %d SASS instructions" % len(ins))
        counts = collections.Counter()
        seen, excerpts = set(), []
        for l in ins:
            body = re.sub(r"^/\*[0-9a-f]+\*/\s+", "", l)
            body = re.sub(r"^@!?U?P\d\s+", "", body)
            op = body.split()[0].rstrip(";")
            if KEEP.match(op):
                counts[op] += 1
            e = EXCERPT.search(body)
            if e and e.group(0) not in seen:
                seen.add(e.group(0))
                excerpts.append(l)
        for op in sorted(counts):
            print("   %-40s x%d" % (op, counts[op]))
        for l in excerpts:
            print("     | " + l)
    return 0


if __name__ == "__main__":
    sys.exit(main())
