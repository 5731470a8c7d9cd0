python -m pytest tests/test_host_layer.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python - <<'PY'
import os, subprocess, sys, tempfile
import numpy as np
sys.path.insert(0, "tests")
import common
from hydrochrono_b200 import h5io
tmp = tempfile.mkdtemp()
h5 = os.path.join(tmp, "sphere.h5")
raw = common.sphere_raw()
h5io.write_bemio(h5, raw)
y = os.path.join(tmp, "c.yaml")
open(y, "w").write("hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: %s\n\n  waves:\n    type: still\n" % h5)
out = os.path.join(tmp, "r.h5")
r = subprocess.run(["hydrochrono_b200/host/build/demo_iea_sphere_yaml", y, out], capture_output=True, text=True)
print(r.returncode, r.stdout[-300:], r.stderr[-300:])
t = h5io.read_f64(out, "results/time/time"); pos = h5io.read_f64(out, "results/model/bodies/body1/position")
g = common.sphere_goldens()
print("HHT rms-rel err", common.rms_relative_error(g["iea_decay_z"], np.interp(g["iea_decay_t"], t, pos[:, 2])))
print("max abs err", np.abs(g["iea_decay_z"] - np.interp(g["iea_decay_t"], t, pos[:, 2])).max())
PY
