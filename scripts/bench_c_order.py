"""A/B of the C-order emit kernels (narrow 256 B row pieces vs wide 1 KiB): PSTB_EMIT_C_NARROW=1 selects the old one."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.argv = [sys.argv[0], "none"]
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import bench_read as br
    br.run("decode+Unit f32 C", 10000, 500000, np.float32, "C", ("unit",))
    br.run("decode only f32 C", 10000, 500000, np.float32, "C", None)
    br.run("decode+Unit f64 C", 10000, 250000, np.float64, "C", ("unit",))
    br.run("decode+Unit f32 C N=50k", 50000, 100000, np.float32, "C", ("unit",))
    br.run("decode+Unit f32 F (reference point)", 10000, 500000, np.float32, "F", ("unit",))
else:
    for env in ("1", "0"):
        print("PSTB_EMIT_C_NARROW=" + env, flush=True)
        subprocess.call([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, PSTB_EMIT_C_NARROW=env))
