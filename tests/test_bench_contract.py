"""CPU: bench.py's reference arm stands on the oracle alone, and the workload definitions both arms use are the same."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_never_touches_the_product():
    """--impl reference renders the bench scene with the oracle only: splat_b200 is never imported and libsplat_b200.so is
    never mapped (the driver lists the .so files a bench process loaded)."""
    code = (
        "import sys, json, io, contextlib\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '3', '--n', '20000']\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf):\n"
        "    args = bench.parse_args()\n"
        "    bench.run_reference(args)\n"
        "maps = open('/proc/self/maps').read()\n"
        "print(json.dumps({'line': json.loads(buf.getvalue()), 'pkg': 'splat_b200' in sys.modules,\n"
        "                  'so': 'libsplat_b200' in maps, 'oracle': 'libsplat_oracle' in maps}))\n"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["pkg"] is False and r["so"] is False and r["oracle"] is True
    line = r["line"]
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["gaussians"] == 20000 and "no scaling" in line["cpu_baseline"]["sample"]


def test_both_arms_share_one_workload_definition(sb, ob):
    sys.path.insert(0, ROOT)
    import bench
    scenes = bench.workload_module()
    assert scenes.GAUSSIAN_DTYPE == sb.GAUSSIAN_DTYPE == ob.GAUSSIAN_DTYPE
    a = scenes.synthetic_gaussians(5000, bench.SCENE_SEED)
    b = sb.scenes.synthetic_gaussians(5000, bench.SCENE_SEED)
    assert a.tobytes() == b.tobytes()
    # the two packers the arms use are byte-identical
    assert np.array_equal(sb.pack_gaussians(b), ob.pack_gaussians(a.view(ob.GAUSSIAN_DTYPE)))
    # SURVEY 8(d) orbit: yaw 2 pi k / 64 + 0.1, pitch 0.1, radius 30
    pos, yaw, pitch = scenes.orbit_camera(16, 64)
    assert abs(yaw - (np.pi / 2 + 0.1)) < 1e-12 and pitch == 0.1 and abs(np.hypot(pos[0], pos[2]) - 30.0) < 1e-9
