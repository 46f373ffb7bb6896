"""Build an alternative libb200tta_<name>.so for A/B runs in ONE gpurun call:
    python scratch/build_variant.py NAME [-DFLAG ...] [attn_fwd.cu=scratch/variants/x.cu.txt ...]
Select it on the box with B200TTA_LIB=longcat_video_tta_b200/libb200tta_NAME.so (see _lib.py)."""
import shutil, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from longcat_video_tta_b200 import _build

name = sys.argv[1]
defs = [a for a in sys.argv[2:] if a.startswith("-D")]
repl = dict(a.split("=") for a in sys.argv[2:] if "=" in a and not a.startswith("-D"))
obj_dir = _build.CSRC / f"build_{name}"
obj_dir.mkdir(exist_ok=True)
src_dir = obj_dir / "src"
src_dir.mkdir(exist_ok=True)
for f in list(_build.CSRC.glob("*.cu")) + list(_build.CSRC.glob("*.cuh")) + list(_build.CSRC.glob("*.h")):
    shutil.copy(f, src_dir / f.name)
for k, v in repl.items():
    shutil.copy(ROOT / v, src_dir / k)
for f in src_dir.iterdir():
    t = f.read_text()
    if "../../include/b200tta.h" in t:
        f.write_text(t.replace("../../include/b200tta.h", "b200tta.h"))


def cc(src):
    obj = obj_dir / (src.stem + ".o")
    cmd = [_build._nvcc(), *_build.NVCC_FLAGS, *defs, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, sorted(src_dir.glob("*.cu"))))
out = _build.PKG / f"libb200tta_{name}.so"
subprocess.run([_build._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *map(str, objs)], check=True)
print(out)
