"""Prints the hot loop (the backward branch around the densest run of LDS.64) of the K1 mean
kernel of a built library: instruction mix per 12 draws (development aid)."""
import re, subprocess, sys, collections

def main(lib, fn="_ZN8b200boot18k1_resample_kernelILi0ELb1EEEvNS_8K1ParamsE", show=False):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    # backward branches
    best = None
    for i, (pc, txt) in enumerate(ins):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < pc:
            tgt = int(m.group(1), 16)
            body = [t for p, t in ins if tgt <= p <= pc]
            n_lds = sum("LDS.64" in t for t in body)
            if n_lds >= 12 and (best is None or len(body) < len(best[3])):
                best = (n_lds, tgt, pc, body)
    n_lds, tgt, pc, body = best
    mix = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in body)
    print(f"{lib}: loop 0x{tgt:x}-0x{pc:x}: {len(body)} instr, {n_lds} LDS.64; total kernel {len(ins)} instr")
    print("   ", dict(mix.most_common()))
    if show:
        for t in body: print("     ", t)

if __name__ == "__main__":
    for lib in sys.argv[1:]:
        if lib == "-v": continue
        main(lib, show="-v" in sys.argv)
