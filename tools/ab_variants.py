"""A/B of compile-time kernel variants in ONE gpurun call (GPU minutes are scarce; a variant costs ~12 s of them).

Workflow
  1. guard the experiment in the .cu source with a macro that defaults to the shipped behaviour
     (`#ifndef LSX_X` / `#define LSX_X 0` / `#endif`);
  2. here (no GPU):   python tools/ab_variants.py build v1="-DLSX_X=1" v2="-DLSX_X=2 -DLSX_Y=1"
     -> langscene-x_b200/csrc/build/variants/liblsx_<name>.so (git-ignored, travels to the GPU box);
     check registers / spills / instruction counts offline first (`-Xptxas -v`, `cuobjdump -sass`);
  3. on the box:      gpurun -- 'python tools/ab_variants.py run C3 > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log'
     times every variant and the default library with tools/stage_times.py (library selected through LSX_B200_LIB),
     then runs the parity tests on the fastest one, falling back to the next if they fail;
  4. hard-wire the winner, delete the macro, and confirm that the default build's SASS equals the tested variant's
     (cuobjdump -sass of the two objects) so that what ships is what was measured.
This is how the last backward-render change of round 1 was found (profiles/r5m_variants.log).
"""
import glob
import json
import os
import shlex
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "langscene-x_b200"))
CSRC = os.path.join(REPO, "langscene-x_b200", "csrc")
OUT = os.path.join(CSRC, "build", "variants")


def build(specs):
    from lsx_b200 import _build
    _build.build_library()  # default objects are reused for every source the flags do not touch
    os.makedirs(OUT, exist_ok=True)
    for spec in specs:
        name, flags = spec.split("=", 1)
        flags = shlex.split(flags)
        macros = [f[2:].split("=")[0] for f in flags if f.startswith("-D")]
        objs = []
        for src in _build.SOURCES:
            path = os.path.join(CSRC, src)
            text = open(path).read() + "".join(open(h).read() for h in _build._headers())
            if any(m in text for m in macros):  # recompile only what can see one of the macros
                obj = os.path.join(OUT, f"{src[:-3]}_{name}.o")
                subprocess.run(["nvcc", *_build.NVCC_FLAGS, *flags, "-c", path, "-o", obj], check=True)
            else:
                obj = os.path.join(_build.OBJ_DIR, src[:-3] + ".o")
            objs.append(obj)
        lib = os.path.join(OUT, f"liblsx_{name}.so")
        subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", lib, "-lcudart"],
                       check=True)
        print("built", lib, " ".join(flags))


def run(configs, stages=("render_fwd", "render_bwd", "preprocess_fwd", "preprocess_bwd")):
    libs = {"default": None}
    for lib in sorted(glob.glob(os.path.join(OUT, "liblsx_*.so"))):
        libs[os.path.basename(lib)[7:-3]] = lib
    results = {}
    for name, lib in libs.items():
        env = dict(os.environ)
        env.pop("LSX_B200_LIB", None)
        if lib:
            env["LSX_B200_LIB"] = lib
        r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "stage_times.py"), *configs], env=env,
                           capture_output=True, text=True, timeout=300)
        rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        if not rows:
            print(f"VARIANT {name}: failed\n{r.stderr[-1500:]}")
            continue
        results[name] = sum(row["sum_ms"] for row in rows)
        for row in rows:
            print(f"VARIANT {name} {row['config']} sum {row['sum_ms']:.4f} ms  " +
                  "  ".join(f"{k} {row['stages_ms'][k]:.4f}" for k in stages), flush=True)
    order = sorted(results, key=results.get)
    print("ORDER", " ".join(order))
    for name in order:
        env = dict(os.environ)
        env.pop("LSX_B200_LIB", None)
        if libs[name]:
            env["LSX_B200_LIB"] = libs[name]
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REPO, "tests", "test_parity_gpu.py"), "-x", "-q",
                            "-m", "gpu"], env=env, capture_output=True, text=True, timeout=600, cwd=REPO)
        print(f"PARITY {name}: rc={r.returncode} {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ''}", flush=True)
        if r.returncode == 0:
            print("WINNER", name)
            return
        print(r.stdout[-2000:])
    print("WINNER none")


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) >= 2 and sys.argv[1] == "run":
        run(sys.argv[2:] or ["C3"])
    else:
        print(__doc__)
