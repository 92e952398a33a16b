"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR) in libb200conformer.so, plus a short
excerpt around the first UTCHMMA of the GEMM kernel.   usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, hashlib, re, subprocess, sys
LIB = "nn_conformer_for_speech_recognition_b200/libb200conformer.so"
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "MUFU", "ACQBULK", "PREEXIT"]
per = collections.OrderedDict(); cur = None; first_gemm = None; lines = sass.splitlines()
for i, l in enumerate(lines):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)[:90]
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for n in names:
            if op.startswith(n):
                per[cur][n] += 1
        if op.startswith("UTCHMMA") and first_gemm is None and "gemm_tcgen05_kernel<256, false, false, 2" in cur:
            first_gemm = (cur, i)
print(f"# {LIB} sha256[:16] = {hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]}; cuobjdump -sass, sm_100a")
print("# kernels that contain tensor-core / TMEM / TMA instructions (counts of SASS mnemonics; HMMA = legacy mma.sync: none expected)")
print(f"{'kernel':92s} " + " ".join(f"{n:>8s}" for n in names) + "    total")
tot = collections.Counter()
for k, c in per.items():
    for n in names:
        tot[n] += c[n]
    if any(c[n] for n in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP")):
        print(f"{k:92s} " + " ".join(f"{c[n]:8d}" for n in names) + f" {c['_total']:8d}")
print(f"{'ALL ' + str(len(per)) + ' kernels':92s} " + " ".join(f"{tot[n]:8d}" for n in names))
if first_gemm:
    k, i = first_gemm
    print(f"\n# excerpt: first UTCHMMA of {k}")
    print("\n".join(lines[max(0, i - 12):i + 10]))
