"""Stage the UNMODIFIED reference sources for the reference arms of bench.py.

    python oracle/stage_reference.py            (called by __graft_entry__.build() where /root/reference exists)

Copies the python / yaml files of the reference's `team_code_v2/` and `lav/` trees byte for byte into the git-ignored
`baseline/_ref/` (it travels to the GPU box with the gpurun snapshot; /root/reference does not exist there).  Nothing is edited:
the two third-party imports the image lacks (torch_scatter, carla) are satisfied by the stand-ins in oracle/refshim at import time.
Test / measurement infrastructure only — nothing under lav_b200/ imports from here."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LAV_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(ref=REF, dst=DST):
    if not os.path.isdir(ref):
        return None
    n = 0
    for top in ("team_code_v2", "lav"):
        for d, _, files in os.walk(os.path.join(ref, top)):
            for f in files:
                if not f.endswith((".py", ".yaml")):
                    continue
                src = os.path.join(d, f)
                out = os.path.join(dst, os.path.relpath(src, ref))
                os.makedirs(os.path.dirname(out), exist_ok=True)
                if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src) or os.path.getsize(out) != os.path.getsize(src):
                    shutil.copy2(src, out)
                n += 1
    for f in ("config_v2.yaml", "config.yaml"):
        if os.path.exists(os.path.join(ref, f)):
            shutil.copy2(os.path.join(ref, f), os.path.join(dst, f))
    return dst, n


if __name__ == "__main__":
    print(stage())
