"""admm_b200 -- host side of the B200-native decentralized TV-ADMM tomography hot path.

Python here is plumbing (device memory via PyTorch, streams, torch.distributed); every floating-point operation
of the path runs in the hand-written sm_100a kernels of libadmm_b200.so (csrc/), called through the C ABI
declared in include/admm_b200.h.  There is no CPU fallback.
"""
from . import _native, registry  # noqa: F401
from .geometry import (angle_split, default_angles_total, graph_csr, make_graph, node_angles, node_to_gpu,  # noqa: F401
                       psnr, shepp_logan, trig_table32)
from .operators import (DenseOperatorCUDA, DensePlan, DiscreteSpace, Element, Plan, RayTransformCUDA,  # noqa: F401
                        make_plan, stack_operators)
