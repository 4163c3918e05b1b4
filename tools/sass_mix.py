"""Instruction mix of a kernel's SASS by issue pipe (cuobjdump, no GPU needed): how many instructions of the traversal
kernels go to the ALU pipe (compares, selects, shifts, logic - half rate), the FMA pipe (fp32 add / mul / fma, integer
multiply-add, moves ptxas places there), memory, control.  usage: sass_mix.py <lib.so> <kernel substring> [...]"""
import collections, re, subprocess, sys

FMA = {"FADD", "FMUL", "FFMA", "IMAD", "DFMA", "DADD", "DMUL", "HFMA2"}
ALU = {"LOP3", "SHF", "SEL", "FSEL", "ISETP", "FSETP", "PLOP3", "IADD3", "LEA", "FMNMX", "VIMNMX", "PRMT", "VIADD", "MOV", "IABS", "FLO", "BREV",
       "POPC", "VIADDMNMX", "FMNMX3", "I2F", "F2I", "CS2R", "BMSK", "SGXT", "LOP", "IADD", "I2FP", "FRND", "F2F", "FCHK", "FSET", "P2R", "R2P", "MUFU"}
MEM = {"LDG", "LD", "LDS", "STG", "ST", "STS", "LDC", "LDCU", "ATOMG", "REDG", "CCTL", "LDL", "STL", "ATOMS"}


def mix(so, name):
    syms = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out, cur = {}, None
    for l in syms.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1) if name in m.group(1) else None
            if cur:
                out[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", l)
        if cur and m:
            op = m.group(1).split()
            op = op[1] if op[0].startswith("@") else op[0]
            b = op.split(".")[0]
            out[cur]["fma" if b in FMA else "alu" if b in ALU else "mem" if b in MEM else "uniform" if b.startswith("U") or b in ("R2UR", "S2UR") else "control"] += 1
    return out


if __name__ == "__main__":
    for name in sys.argv[2:]:
        for k, c in mix(sys.argv[1], name).items():
            print(f"{k[:60]:60s} total {sum(c.values()):5d}  alu {c['alu']:4d}  fma {c['fma']:4d}  mem {c['mem']:4d}  control {c['control']:4d}  uniform {c['uniform']:4d}")
