"""Drop-in for the reference's two `load_odl_data` variants: the operator-building one
(block_2_load_odl_data.py:99-253) and the pickle-loading one every block_7_main_ver* calls
(block_2_test.py:15-167).  Operators are built in-process as matrix-free CUDA ray transforms -- no ODL, no
pickles, no dense matrices; the returned dict carries the keys both variants return."""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

from admm_b200 import RayTransformCUDA, angle_split, default_angles_total, node_angles, registry, stack_operators
from Gen_Sino_Partitioned import ConstIm, randIm  # noqa: F401


def _build_parallel_beam_operators(N, num_nodes, angles_total=None, det_width_factor=1.0, partition="reference_literal",
                                   device=0):
    """block_2_load_odl_data.py:16-65.  Image space [-1,1]^2, N x N, float32 (:23-28); angles_total default
    max(180, 3N) (:31-33); integer split (:36-38); detector width det_width_factor*2 with N bins (:42-44).
    `partition`: "reference_literal" (default) is what the shipped code does -- every node gets the
    midpoints of uniform_partition(0, pi, m_k) (:51), the aggregate its own pi/angles_total grid (:59-63);
    "contiguous" gives node k the k-th block of the aggregate grid (north_star's angle-partitioned operators, what
    bench.py times; then the aggregate is the vstack of the nodes)."""
    if angles_total is None:
        angles_total = default_angles_total(N)
    det_width = det_width_factor * 2.0
    thetas = node_angles(angles_total, num_nodes, partition)
    assert [len(t) for t in thetas] == angle_split(angles_total, num_nodes)
    cell = np.pi / angles_total if partition == "contiguous" else None       # else pi / m_k (RayTransformCUDA default)
    ray_transforms = [RayTransformCUDA(N, t, D=N, det_w=det_width, device=device, angle_cell=cell) for t in thetas]
    if partition == "contiguous":
        agg = stack_operators(ray_transforms)
    else:
        agg = RayTransformCUDA(N, (np.arange(angles_total) + 0.5) * np.pi / angles_total, D=N, det_w=det_width,
                               device=device)
    return ray_transforms[0].domain, ray_transforms, agg


def load_odl_data(base_dir="saved_operators_Incmp_Span", N=64, num_nodes=5, noise_level=0.005, phantom_array=None,
                  ray_transforms_pickle=None, A_dense_list_pickle=None, agg_op_pickle=None, A_agg_pickle=None,
                  make_plots=False, show_plots=False, output_dir=None, save_operators_dir=None, build_dense=False,
                  angles_total=None, partition="reference_literal", seed=None, device=0):
    """Union of both reference signatures (block_2_test.py:15-26, block_2_load_odl_data.py:99-109).

    Sinograms: b_i = A_i x + noise_level * N(0, 1) (block_2_load_odl_data.py:148-154 / block_2_test.py:54-60), drawn
    from the global np.random like the reference unless `seed` is given.  A single phantom is shared by all nodes
    (block_2_test.py:48-51) unless `phantom_array` is a list of per-node phantoms (block_2_load_odl_data.py:139-143).
    Pickle arguments are accepted and ignored; `build_dense` / plotting are out of scope."""
    space, ray_transforms, agg_ray_trafo = _build_parallel_beam_operators(N, num_nodes, angles_total,
                                                                         partition=partition, device=device)
    rng = np.random if seed is None else np.random.RandomState(seed)
    if phantom_array is None:
        phantom_array = randIm(N) if seed is None else randIm(N, seed=seed)
    if isinstance(phantom_array, list):
        assert len(phantom_array) == num_nodes, "phantom_array list must have length num_nodes"
        phantoms = [np.asarray(p) for p in phantom_array]
    else:
        phantoms = [np.asarray(phantom_array) for _ in range(num_nodes)]
    sinograms = []
    for i, op in enumerate(ray_transforms):
        clean = op(space.element(phantoms[i]))
        noisy = clean + noise_level * op.range.element(rng.normal(0.0, 1.0, size=op.range.shape))
        sinograms.append(noisy.asarray())
    agg_sinogram = np.vstack(sinograms)                      # block_2_test.py:65-66
    Wi_list = [np.maximum(op.colnorm2(), 1e-12) for op in ray_transforms]   # block_3...:22-23 (data_block2.pkl key)
    column_norms_all = [np.sqrt(w) for w in Wi_list]                        # np.linalg.norm(A_i, axis=0), :62
    if output_dir is None:
        output_dir = f"Recon_Op_ADMM_{datetime.now().strftime('%Y%m%d_%H%M%S')}"
    for key in (base_dir, save_operators_dir):      # what the reference pickles into these directories
        if key is not None:
            registry.put(key, ray_transforms)
    return {
        "A_dense_list": ray_transforms,          # matrix-free: .shape, @, .T, .colnorm2()
        "ray_transforms": ray_transforms,
        "sinograms": sinograms,
        "column_norms_all": column_norms_all,
        "Wi_list": Wi_list,
        "N": N,
        "num_nodes": num_nodes,
        "agg_ray_trafo": agg_ray_trafo,
        "A_agg": agg_ray_trafo,
        "agg_sinogram": agg_sinogram,
        "agg_fbp_recon": None,
        "agg_ls_recon": None,
        "output_dir": output_dir,
        "phantom": np.asarray(phantoms[0], dtype=np.float32),
        "phantoms": [np.asarray(p, dtype=np.float32) for p in phantoms],
    }


load_data = load_odl_data
prepare_data = load_odl_data
