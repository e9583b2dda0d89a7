"""Count the SASS mnemonics that identify the Blackwell features per kernel.
usage: cuobjdump -sass breaching_b200/lib/libbreaching_b200.so | python profiles/sass_summary.py > profiles/sass_summary_rN.txt"""
import collections
import re
import sys

txt = sys.stdin.read()
funcs = re.split(r"\n\s*Function : ", txt)
PAT = {
    "UTCHMMA (tcgen05.mma)": r"\bUTCHMMA", "LDTM (tcgen05.ld)": r"\bLDTM", "UTCBAR (tcgen05.commit)": r"\bUTCBAR",
    "UTCATOMSWS (tcgen05.alloc / dealloc)": r"UTCATOMSWS", "UTMALDG.4D.IM2COL (TMA im2col)": r"UTMALDG\.\dD\.IM2COL",
    "UTMALDG.2D/3D (TMA tile)": r"UTMALDG\.\dD(?!\.IM2COL)", "SYNCS (mbarrier)": r"\bSYNCS",
    "UCGABAR (cluster barrier)": r"UCGABAR_ARV", "LDGSTS (cp.async)": r"\bLDGSTS", "LDG.E.128": r"LDG\.E\.128", "STG.E.128": r"STG\.E\.128",
    "ACQBULK / griddepcontrol (PDL)": r"ACQBULK",
}
tot, per = collections.Counter(), collections.defaultdict(collections.Counter)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    m = re.search(r"(\d+)([a-z_0-9]+_kernel)(I[A-Za-z0-9]*E)?", name)
    short = (m.group(2) + (m.group(3) or "")) if m else name[:80]
    for k, p in PAT.items():
        n = len(re.findall(p, f))
        if n:
            tot[k] += n
            per[k][short] += n
print("# SASS mnemonic counts of breaching_b200/lib/libbreaching_b200.so (cuobjdump -sass, sm_100a)")
print("# template suffix of igemm_tc_kernel: I Li<mode 0 fprop / 1 dgrad / 2 wgrad> Li<BN> Lb<TMA producer> Lb<per-class strided dgrad> E")
for k, v in tot.items():
    print(f"{v:6d}  {k}")
    for fn, n in per[k].most_common(12):
        print(f"          {n:5d}  {fn}")
