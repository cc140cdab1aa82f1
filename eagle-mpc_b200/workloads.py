"""Named workloads of BASELINE.json (synthetic batches of the reference's YAML problems)."""
import numpy as np

CONFIGS = {
    # name: (yaml, dt_ms, seed0)
    "hexacopter370_hover": ("hexacopter370/trajectories/hover.yaml", 20, 1000),
    "hexacopter370_passthrough": ("hexacopter370/trajectories/passthrough.yaml", 20, 1500),
    "hexacopter370_flying_arm_3_displacement": ("hexacopter370_flying_arm_3/trajectories/displacement.yaml", 20, 2024),
    "hextilt_flying_arm_5_push_slide": ("hextilt_flying_arm_5/trajectories/push_slide.yaml", 20, 4096),
    "iris_px4_displacement": ("iris_px4/trajectories/displacement.yaml", 20, 9000),
    "iris_px4_hover": ("iris_px4/trajectories/hover.yaml", 20, 9500),
}


def noisy_x0(x0, batch, seed0, first=0):
    """benchmark/utils/utils.hpp:15-27 recipe: x += 0.05*U(-1,1)^nx elementwise, then renormalise the quaternion
    x[3:7].  OCP b uses numpy's MT19937 generator seeded with seed0 + b (the reference used an unseeded
    Eigen::VectorXd::Random)."""
    x0 = np.asarray(x0, dtype=np.float64)
    out = np.empty((batch, x0.size))
    for i in range(batch):
        rng = np.random.Generator(np.random.MT19937(seed0 + first + i))
        x = x0 + 0.05 * rng.uniform(-1.0, 1.0, size=x0.size)
        x[3:7] /= np.linalg.norm(x[3:7])
        out[i] = x
    return out


def algorithmic_bytes_per_node(nx, ndx, nu, n_trials=1):
    """algorithmic HBM bytes per node-iteration."""
    D = 2 * ndx * ndx + 2 * ndx * nu + nu * nu + ndx + nu
    return (8 * ((nx + nu) + D + nx + 1) + 8 * ((D + ndx) + (nu * ndx + nu + ndx + ndx * ndx)) +
            8 * n_trials * ((nx + nu + nu * ndx + nu + ndx) + (nx + nu + 1)))
