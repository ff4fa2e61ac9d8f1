"""Builds alternative libb200clip variants for same-box A/B timing (run them with B200CLIP_LIB=<path>).
usage: build_variants.py name:-DFLAG=V[,-DFLAG2=V2][@git-rev-of-gemm.cu] ...
Each variant recompiles only csrc/gemm.cu (optionally taken from a git revision) and links it with the
objects of the regular build into construction_clip_b200/_variants/libb200clip_<name>.so."""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from construction_clip_b200 import build as B  # noqa: E402

B.build()
out_dir = B.PKG / "_variants"
out_dir.mkdir(exist_ok=True)
nvcc = B._nvcc()
for spec in sys.argv[1:]:
    name, _, rest = spec.partition(":")
    flags, _, rev = rest.partition("@")
    src = B.CSRC / "gemm.cu"
    if rev:
        text = subprocess.run(["git", "show", f"{rev}:construction_clip_b200/csrc/gemm.cu"], capture_output=True,
                              text=True, check=True, cwd=B.PKG.parent).stdout
        src = out_dir / f"gemm_{name}.cu"
        src.write_text(text)
    obj = out_dir / f"gemm_{name}.o"
    cmd = [nvcc, *B.NVCC_FLAGS, f"-I{B.CSRC}", *[f for f in flags.split(",") if f], "-c", str(src), "-o", str(obj)]
    subprocess.run(cmd, check=True)
    objs = [str(o) for o in sorted(B.OBJ.glob("*.o")) if o.name != "gemm.o"] + [str(obj)]
    lib = out_dir / f"libb200clip_{name}.so"
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *objs], check=True)
    print(lib)
