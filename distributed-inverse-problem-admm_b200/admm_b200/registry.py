"""Process-local hand-over of the node operators between the drop-in modules.

The reference's drivers pass a directory name instead of the operators: `load_odl_data(base_dir=...)` reads pickles
from it and `build_pixel_connected_Q_provider(base_dir=...)` reads `A_dense_list.pkl` from it again
(block_2_test.py:28-31, block_3_graph_and_precisions.py:288-291).  Here nothing is pickled (the operators are
matrix-free CUDA objects), so block_2 registers what it built under the same key and block_3 looks it up."""
_OPS = {}


def put(key, ops):
    _OPS[str(key)] = list(ops)


def get(key):
    return _OPS.get(str(key))
