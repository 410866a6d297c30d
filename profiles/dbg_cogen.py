"""Cogen plant-model accuracy of the tensor-core (3xTF32) and FFMA kernels vs the f32/f64 NumPy evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "cogen_model_io.npz"))
z = np.load(os.path.join(ROOT, "sustaingym_b200", "data", "cogen_plant.npz"))
span = z["out_max"] - z["out_min"]; scale = np.where(np.isfinite(span) & (span > 0), span, 100.0)[None]
for tcf in (True, False):
    env = BatchedCogenEnv(1, tensor_cores=tcf)
    Y = env.plant(torch.from_numpy(g["X"]).cuda()).cpu().numpy()
    e32 = np.abs(Y - g["Y32"]) / np.maximum(np.abs(g["Y32"]), scale)
    e64 = np.abs(Y - g["Y64"]) / np.maximum(np.abs(g["Y64"]), scale)
    print(f"tensor_cores={tcf}: max scaled err vs f32 numpy {e32.max():.3e} (p99 {np.percentile(e32, 99):.2e}), vs f64 {e64.max():.3e}; "
          f"abs err max {np.abs(Y - g['Y64']).max():.3e}; worst column {int(np.unravel_index(e64.argmax(), e64.shape)[1])}")
ref = np.abs(g["Y32"] - g["Y64"]) / np.maximum(np.abs(g["Y64"]), scale)
print(f"f32 numpy vs f64 numpy: {ref.max():.3e}")
