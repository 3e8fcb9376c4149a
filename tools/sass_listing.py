"""Disassemble the three hot kernels of libbosip_b200.so (cuobjdump -sass) into profiles/sass_<kernel>.txt: a mnemonic
histogram (the tcgen05 / TMA / DMMA evidence: UTCIMMA, UTCBAR, LDTM, STTM, UBLKCP, DMMA ...) followed by the full listing.
    python tools/sass_listing.py            (no GPU needed)"""
import collections, os, re, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "bosip.jl_b200", "libbosip_b200.so")
OUT = os.path.join(ROOT, "profiles")
# output name -> (substring of the mangled name, further required substring)
KERNELS = {
    "vargemm_i8": ("vargemm_i8_kernelILi7ELi2ELi0ELi6E", ""),          # default engine: 7 slices of alpha^2 L^-1 x 6 of K*, two MMA issuers, plain schedule
    "kstar": ("kstar_kernelILi2ELi5ELi2ELi6E", ""),                    # Matern-5/2, d = 5, INT8 slice mode, 6 slices of K* (config C3)
    "vargemm_i8_pair": ("vargemm_i8p_kernel", ""),                     # opt-in paired kernel (tcgen05 cta_group::2, multicast commits)
    "vargemm": ("14vargemm_kernel", ""),                               # FP64 DMMA engine
}
KEY = ("UTCIMMA", "UTCBAR", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "DMMA", "SYNCS", "DFMA", "DADD", "DMUL", "IMAD.WIDE", "PRMT", "MUFU", "F2I", "I2F",
       "LDS", "STS", "LDG", "STG", "BAR", "ELECT", "R2UR")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"(?m)^\s*Function : ", txt)[1:]
    for out_name, (needle, needle2) in KERNELS.items():
        hits = [b for b in blocks if needle in b.split("\n", 1)[0] and needle2 in b.split("\n", 1)[0]]
        if not hits:
            print(f"{out_name}: no function matches {needle}", file=sys.stderr)
            continue
        body = hits[0]
        name = body.split("\n", 1)[0].strip()
        ops = collections.Counter()
        n_instr = 0
        for line in body.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                n_instr += 1
                ops[m.group(1)] += 1
        with open(os.path.join(OUT, f"sass_{out_name}.txt"), "w") as f:
            f.write(f"# cuobjdump -sass bosip.jl_b200/libbosip_b200.so  (sm_100a), function {name}\n")
            f.write(f"# {n_instr} SASS instructions.  Counts of the instructions that identify the engine:\n")
            for k in KEY:
                c = sum(v for op, v in ops.items() if op.startswith(k))
                if c:
                    f.write(f"#   {k:10s} {c}\n")
            f.write("# full opcode histogram: " + ", ".join(f"{op} {c}" for op, c in ops.most_common()) + "\n\n")
            f.write("        Function : " + name + "\n")
            for line in body.splitlines()[1:]:
                line = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", line)      # drop the encoding words
                if line.strip():
                    f.write(line.rstrip() + "\n")
        print(out_name, name, n_instr, {k: sum(v for op, v in ops.items() if op.startswith(k)) for k in KEY[:8]})


if __name__ == "__main__":
    main()
