"""SASS of the two render kernels as shipped (headline instantiation: 28 channels), for checking instruction-count claims:
    python tools/sass_extract.py        -> profiles/sass_render_fwd.txt, profiles/sass_render_bwd.txt
Each file: the function's SASS without the encoding columns, then a histogram of its opcodes and the mnemonics that matter
for the design claims (LDGSTS = per-lane async copies, RED = global reductions, UTMA*/UBLKCP/UTC*MMA = none: no TMA, no tensor cores)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(REPO, "langscene-x_b200", "csrc", "build")


def function_sass(obj, pattern):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = re.search(pattern, line) is not None
            if on:
                out.append(line.strip())
            continue
        if on:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append(f"{m.group(1)}  {m.group(2).strip()}")
    return out


def main():
    for name, obj, pat in (("fwd", "render_fwd.o", r"render_fwd_kernelILi28E"), ("bwd", "render_bwd.o", r"render_bwd_kernelILi28ELi20E")):
        lines = function_sass(os.path.join(OBJ, obj), pat)
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", l.split("  ", 1)[1]).split()[0].split(".")[0] for l in lines[1:])
        path = os.path.join(REPO, "profiles", f"sass_render_{name}.txt")
        with open(path, "w") as f:
            f.write(f"# cuobjdump -sass of {obj} ({len(lines) - 1} instructions), nvcc 12.9, sm_100a; regenerate: python tools/sass_extract.py\n")
            f.write("\n".join(lines) + "\n\n# opcode histogram\n")
            for op, n in ops.most_common():
                f.write(f"{op:12s} {n}\n")
            f.write("\n# design-claim mnemonics: " + ", ".join(f"{k}={sum(v for o, v in ops.items() if o.startswith(k))}"
                                                             for k in ("LDGSTS", "RED", "ATOM", "UTMA", "UBLKCP", "UTC", "LDTM", "HMMA", "MUFU", "SHFL", "LDS", "STS")) + "\n")
        print(path, len(lines) - 1, "instructions")


if __name__ == "__main__":
    sys.exit(main())
