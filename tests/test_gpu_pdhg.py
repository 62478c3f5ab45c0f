"""GPU parity of the PDHG consensus variant (ADMM_Tomo_Only.py:89-148) against its fp64 oracle twin:
same operators, data, phantom and step-size rule on both sides, through the C ABI (admm_pdhg_* + K1 / K2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,M,V,partition", [(32, 60, 3, "contiguous"), (48, 45, 5, "reference_literal")])
def test_pdhg_consensus_matches_oracle(N, M, V, partition):
    from admm_b200 import RayTransformCUDA, node_angles
    from ADMM_Tomo_Only import pdhg_consensus
    from oracle import oracle as O
    thetas = node_angles(M, V, partition)
    img = O.shepp_logan(N)
    ops_o = [O.JosephOperator(N, t) for t in thetas]
    sinos = [(op.forward(img) + 0.005 * np.random.default_rng(7 + i).standard_normal(op.shape[0])).astype(np.float32)
             for i, op in enumerate(ops_o)]
    ops_g = [RayTransformCUDA(N, t) for t in thetas]
    kw = dict(niter=8, lambda_penalty=0.005, lambda_agg=0.005, gamma=2.0, node_niter=5, agg_niter=15, opnorm_iters=20)
    ro = O.pdhg_consensus(ops_o, sinos, img, N, **kw)
    rg = pdhg_consensus(ops_g, sinos, img, **kw)
    err = {"opnorm": max(abs(a - b) / b for a, b in zip(rg["op_norms"] + [rg["op_norm_agg"]], ro["op_norms"] + [ro["op_norm_agg"]]))}
    for key in ("mse_lists", "mse_sino_lists"):
        err[key] = float(np.max(np.abs(np.array(rg[key]) - np.array(ro[key])) / np.array(ro[key])))
    for key in ("mse_agg_list", "mse_agg_sino_list"):
        err[key] = float(np.max(np.abs(np.array(rg[key]) - np.array(ro[key])) / np.array(ro[key])))
    err["x_agg"] = float(np.linalg.norm(rg["x_agg"].reshape(-1) - ro["x_agg"]) / np.linalg.norm(ro["x_agg"]))
    err["x_vars"] = float(max(np.linalg.norm(a.reshape(-1) - b) / max(np.linalg.norm(b), 1e-30)
                              for a, b in zip(rg["x_vars"], ro["x_vars"])))
    print("PARITY pdhg " + ", ".join(f"{k} {v:.2e}" for k, v in err.items()))
    assert err["opnorm"] < 1e-4, err
    for key in ("mse_lists", "mse_sino_lists", "mse_agg_list", "mse_agg_sino_list", "x_agg", "x_vars"):
        assert err[key] < 1e-3, (key, err)
    # the aggregate problem makes progress (the node problems are held near their start by the consensus pull)
    assert rg["mse_agg_list"][-1] < rg["mse_agg_list"][0]
