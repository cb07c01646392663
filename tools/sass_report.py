"""-Xptxas -v resource lines and a SASS opcode histogram for the hot kernels of libnzcb.so's objects (CPU only:
cuobjdump on nzcb_circom_b200/build/*.o).   python tools/sass_report.py > profiles/rNN_ptxas_sass_hot_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = [("msm.o", "k_aff_backwardILb0"), ("msm.o", "k_aff_backwardILb1"), ("msm.o", "k_aff_forwardILb1"), ("msm.o", "k_aff_invert"),
       ("msm.o", "k_msm_accumILb1"), ("msm.o", "k_msm_accumILb0"), ("msm.o", "k_msm_digitsILb0ELb0"), ("ntt.o", "k_ntt_passILi3ELi256ELi2"),
       ("plonk.o", "k_round3"), ("witness.o", "k_witnessILj384ELj1"), ("witness.o", "k_witnessILj128ELj3"),
       ("witness.o", "sha_step_warpIjLj6145"), ("witness.o", "quinsel_warpILj6145"), ("verify.o", "k_plonk_verify")]


def main():
    for obj, pat in HOT:
        path = os.path.join(ROOT, "nzcb_circom_b200", "build", obj)
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        # resource usage: "Function NAME:" followed by " REG:.. STACK:.."
        m = re.search(r"Function (\S*" + re.escape(pat) + r"\S*):\s*\n\s*(REG:[^\n]*)", res)
        print(f"== {obj}: {pat}")
        if m:
            print("   " + m.group(2).strip())
        blocks = re.split(r"\n\s*Function : ", sass)
        body = next((b for b in blocks if pat in b.split("\n", 1)[0]), None)
        if body is None:
            print("   (not found in SASS)")
            continue
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", body, flags=re.M)
        h = collections.Counter(ops)
        total = sum(h.values())
        fma = sum(v for k, v in h.items() if k.startswith(("IMAD", "FFMA", "FMUL", "HFMA")))
        wide = sum(v for k, v in h.items() if k.startswith("IMAD.WIDE"))
        print(f"   {total} SASS instructions; multiplier pipe {fma} ({100 * fma / total:.0f} %), of which IMAD.WIDE {wide}; "
              f"tensor / TMA instructions: {sum(v for k, v in h.items() if k.startswith(('UTC', 'HMMA', 'UTMA', 'UBLKCP', 'LDTM')))}")
        print("   " + ", ".join(f"{k} {v}" for k, v in h.most_common(14)))


if __name__ == "__main__":
    main()
