"""Register use per `setmaxnreg` region of the ping-pong forward kernels, from the SASS of the built object:

    python tools/sass_regions.py [scade_b200/_lib/mlp_tc.o]

For every nerf_mlp_tc_pp_kernel instantiation: the highest register index between the two USETMAXREG instructions (the control
warps: TMA producer, MMA issuer, compositor -- they keep 32 registers) and inside every local call target, plus spill counts.
A control-region index above R31 would be a bug: the warp no longer owns those registers."""
import re
import subprocess
import sys

obj = sys.argv[1] if len(sys.argv) > 1 else "scade_b200/_lib/mlp_tc.o"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0]
    if "nerf_mlp_tc_pp_kernel" not in name:
        continue
    lines = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", part)]
    regs = lambda seg: max((int(x) for ins in seg for x in re.findall(r"\bR(\d+)\b", ins)), default=-1)
    sm = [a for a, i in lines if "USETMAXREG" in i]
    print(name[:100])
    print(f"  whole kernel: max R{regs([i for _, i in lines])}")
    if len(sm) == 2:
        print(f"  control region [{sm[0]:#x}, {sm[1]:#x}): max R{regs([i for a, i in lines if sm[0] <= a < sm[1]])}")
    targets = sorted(set(int(x, 16) for x in re.findall(r"CALL\.REL\.NOINC 0x([0-9a-f]+)", part)))
    for k, t in enumerate(targets):
        end = targets[k + 1] if k + 1 < len(targets) else lines[-1][0] + 16
        seg = [i for a, i in lines if t <= a < end]
        if len(seg) > 40:
            print(f"  call target {t:#x}: {len(seg)} instructions, max R{regs(seg)}, LDL {sum('LDL' in i for i in seg)}, STL {sum('STL' in i for i in seg)}")
