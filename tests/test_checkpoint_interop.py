"""Checkpoint interop: the reference's own trained tf.train.Checkpoint (quadcopter, 2 experts x 2 outputs,
M = 100; copied as data to tests/golden/quadcopter_ckpt/) is read without TensorFlow, loaded into the legacy-flavour
model, and evaluated on real quadcopter rows against tests/golden/quadcopter_golden.npz (oracle outputs at the
reference-trained parameters; generator: tests/golden/make_quadcopter_golden.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT_DIR = os.path.join(HERE, "golden", "quadcopter_ckpt")
GOLD = os.path.join(HERE, "golden", "quadcopter_golden.npz")


def _model(bound="further_gating"):
    from mogpe_b200 import gpflow_lite as gl
    from mogpe_b200.experts import SVGPExpert, SVGPExperts
    from mogpe_b200.gating_networks import SVGPGatingNetwork
    from mogpe_b200.mixture_of_experts import MixtureOfSVGPExperts

    D, F, M = 2, 2, 100

    def expert():
        return SVGPExpert(gl.SeparateIndependent([gl.RBF(lengthscales=np.ones(D)) for _ in range(F)]), gl.Gaussian(np.ones((1, F))),
                          gl.SharedIndependentInducingVariables(gl.InducingPoints(np.zeros((M, D)))), gl.Constant(np.zeros(F)),
                          num_latent_gps=F)

    gate = SVGPGatingNetwork(gl.SeparateIndependent([gl.RBF(lengthscales=np.ones(D))]), gl.Bernoulli(),
                             gl.InducingPoints(np.zeros((M, D))), None)
    return MixtureOfSVGPExperts(gate, SVGPExperts([expert(), expert()]), num_data=2685, num_samples=1, bound=bound)


def _eps(g):
    return {"u_experts": list(g["eps/u_experts"]), "u_gating": g["eps/u_gating"], "h": g["eps/h"], "f": g["eps/f"]}


def test_reader_parses_the_reference_checkpoint():
    from mogpe_b200.training.tf_checkpoint import read_index, read_tf_checkpoint

    t = read_tf_checkpoint(os.path.join(CKPT_DIR, "ckpt-200"))
    assert len(t) == 24 and int(t["save_counter/.ATTRIBUTES/VARIABLE_VALUE"]) == 200
    q = t["model/experts/experts_list/0/q_sqrt/_pretransformed_input/.ATTRIBUTES/VARIABLE_VALUE"]
    assert q.shape == (2, 5050) and q.dtype == np.float64
    ls = t["model/experts/experts_list/0/kernel/kernels/1/lengthscales/_pretransformed_input/.ATTRIBUTES/VARIABLE_VALUE"]
    assert ls[0] == pytest.approx(97.60962219, rel=1e-9)
    idx = read_index(os.path.join(CKPT_DIR, "ckpt-200"))
    assert "_CHECKPOINTABLE_OBJECT_GRAPH" in idx  # string tensor: listed, skipped by the numeric reader
    with pytest.raises(ValueError, match="bad magic"):
        bad = os.path.join(CKPT_DIR, "ckpt-200.data-00000-of-00001")[:-len(".data-00000-of-00001")]
        import shutil, tempfile
        d = tempfile.mkdtemp()
        shutil.copy(os.path.join(CKPT_DIR, "ckpt-200.data-00000-of-00001"), os.path.join(d, "x.index"))
        read_index(os.path.join(d, "x"))


def test_checkpoint_loads_into_the_legacy_model():
    from mogpe_b200.training.utils import assign_checkpoint_tensors, update_model_from_checkpoint

    model = update_model_from_checkpoint(_model(), CKPT_DIR)
    e0 = model.experts.experts_list[0]
    assert e0.kernel.kernels[1].lengthscales.numpy()[0] == pytest.approx(97.60962219, rel=1e-9)  # softplus(97.6) = 97.6
    assert e0.likelihood.variance.numpy().shape == (1, 2) and e0.likelihood.variance.numpy().min() > 1e-6
    assert not np.any(np.triu(e0.q_sqrt.numpy()[0], 1)) and e0.q_sqrt.numpy().shape == (2, 100, 100)
    assert float(model.gating_network.kernel.kernels[0].variance.numpy()) == pytest.approx(3.3173931644, rel=1e-9)
    with pytest.raises(ValueError, match="without a counterpart"):
        assign_checkpoint_tensors(model, {"model/experts/experts_list/2/q_mu": np.zeros((100, 2))})
    with pytest.raises(ValueError, match="checkpoint has shape"):
        assign_checkpoint_tensors(model, {"model/gating_network/q_mu": np.zeros((50, 1))})


def _spec_from_model(model, bound):
    def gp(m, expert):
        ks = m.kernel.kernels
        d = {"Z": m.Z.numpy(), "kernels": [{"lengthscales": k.lengthscales.numpy(), "variance": float(k.variance.numpy())} for k in ks],
             "q_mu": m.q_mu.numpy(), "q_sqrt": m.q_sqrt.numpy(), "mean_c": m.mean_function.c.numpy() if expert else None}
        if expert:
            d["noise_variance"] = m.likelihood.variance.numpy().reshape(-1)
        return d

    return {"flavour": "legacy", "bound": bound, "num_samples": 1, "num_data": 2685,
            "experts": [gp(e, True) for e in model.experts.experts_list], "gating": gp(model.gating_network, False)}


def test_oracle_at_trained_parameters_reproduces_golden():
    from mogpe_b200.training.utils import update_model_from_checkpoint
    from oracle import ref_numpy as rn

    g = np.load(GOLD)
    model = update_model_from_checkpoint(_model(), CKPT_DIR)
    for b in ("further_gating", "tight", "further"):
        v = rn.elbo(_spec_from_model(model, b), g["X"], g["Y"], _eps(g))
        assert v == pytest.approx(float(g[f"elbo/{b}"]), rel=1e-11)


@pytest.mark.gpu
def test_cuda_path_at_reference_trained_parameters():
    from mogpe_b200.training.utils import update_model_from_checkpoint

    g = np.load(GOLD)
    X, Y, eps = g["X"], g["Y"], _eps(g)
    for b in ("further_gating", "tight", "further"):
        model = update_model_from_checkpoint(_model(b), CKPT_DIR)
        v = model.maximum_log_likelihood_objective((X, Y), eps=eps)
        assert abs(v - float(g[f"elbo/{b}"])) <= 1e-8 * abs(float(g[f"elbo/{b}"])), (b, v, float(g[f"elbo/{b}"]))
    dist = model.predict_y(X)
    assert np.max(np.abs(dist.mean() - g["y_mean"])) <= 1e-8 * np.max(np.abs(g["y_mean"]))
    assert np.max(np.abs(dist.variance() - g["y_var"])) <= 1e-8 * np.max(np.abs(g["y_var"]))
    assert np.max(np.abs(model.predict_mixing_probs(X) - g["mix_probs"])) <= 1e-8
    # the reference's metrics on top of predict_y (mogpe/training/metrics.py) at trained parameters: sane values
    from mogpe_b200.training import metrics

    nlpd, mae, rmse = (f(model, (X, Y)) for f in (metrics.negative_log_predictive_density, metrics.mean_absolute_error,
                                                  metrics.root_mean_squared_error))
    assert mae == pytest.approx(float(np.mean(np.abs(g["y_mean"] - Y))), rel=1e-8)
    assert rmse == pytest.approx(float(np.sqrt(np.mean((g["y_mean"] - Y) ** 2))), rel=1e-8)
    assert np.isfinite(nlpd) and nlpd < 0.5 and rmse < 0.5  # a trained model fits standardised data well
    fm, fv = model.experts.experts_list[0].predict_f(X)
    assert np.max(np.abs(fm - g["expert0/f_mean"])) <= 1e-8 * np.max(np.abs(g["expert0/f_mean"]))
    assert np.max(np.abs(fv - g["expert0/f_var"])) <= 1e-8 * np.max(np.abs(g["expert0/f_var"]))


def test_writer_primitives_are_pinned_by_the_reference_checkpoint():
    """CRC-32C + masking, the string-tensor encoding and the block trailer of the writer reproduce the bytes of the
    reference's own checkpoint; the object graph generated from the variable keys places every variable on the same path of
    child names as TensorFlow's graph does."""
    import struct

    from mogpe_b200.training import tf_checkpoint as tc

    pre = os.path.join(CKPT_DIR, "ckpt-200")
    buf = open(pre + ".index", "rb").read()
    data = open(pre + ".data-00000-of-00001", "rb").read()
    assert tc.crc32c(b"123456789") == 0xE3069283  # the CRC-32C check value
    footer = buf[-48:]
    _, p = tc._varint(footer, 0)
    _, p = tc._varint(footer, p)
    ioff, p = tc._varint(footer, p)
    isize, p = tc._varint(footer, p)
    graph_payload, n_tensors = None, 0
    for _, handle in tc._block_entries(buf, ioff, isize):
        boff, q = tc._varint(handle, 0)
        bsize, q = tc._varint(handle, q)
        block, trailer = buf[boff:boff + bsize], buf[boff + bsize:boff + bsize + 5]
        assert struct.unpack("<I", trailer[1:])[0] == tc.mask_crc(tc.crc32c(block + trailer[:1]))
        for key, value in tc._block_entries(buf, boff, bsize):
            if key == b"":
                assert value == tc._pb_varint(1, 1) + tc._pb_bytes(3, tc._pb_varint(1, 1))  # BundleHeaderProto
                continue
            m = tc._parse_proto(value)
            raw = data[m.get(4, [0])[0]:m.get(4, [0])[0] + m[5][0]]
            if m[1][0] == 7:
                ln, pp = tc._varint(raw, 0)
                graph_payload = raw[pp + 4:]
                enc, crc = tc.encode_string_tensor(graph_payload)
                assert enc == raw and crc == m[6][0]
            else:
                assert tc.mask_crc(tc.crc32c(raw)) == m[6][0]
                shape = tc._shape(m[2][0]) if 2 in m else ()
                assert value == tc._entry_proto(m[1][0], shape, m.get(4, [0])[0], m[5][0], m[6][0])  # same message bytes
                n_tensors += 1
    assert n_tensors == 24
    theirs = tc.parse_object_graph(graph_payload)
    keys = [k for k in tc.read_index(pre) if k != "_CHECKPOINTABLE_OBJECT_GRAPH"]
    mine = tc.parse_object_graph(tc.object_graph_proto(keys))
    assert mine == theirs and len(mine) == 24


def test_checkpoint_round_trip_through_the_writer(tmp_path):
    """Reference checkpoint -> model -> CheckpointManager.save -> fresh model: identical variables; the written files parse
    as a tensor bundle with the reference's variable names, and max_to_keep prunes old files."""
    from mogpe_b200.training.tf_checkpoint import read_index, read_tf_checkpoint
    from mogpe_b200.training.utils import init_checkpoint_manager, update_model_from_checkpoint

    a = update_model_from_checkpoint(_model(), CKPT_DIR)
    mgr = init_checkpoint_manager(a, str(tmp_path / "log"), num_ckpts=2)
    for _ in range(3):
        prefix = mgr.save()
    assert os.path.basename(prefix) == "ckpt-3" and not os.path.exists(str(tmp_path / "log" / "ckpt-1.index"))
    ref = read_tf_checkpoint(os.path.join(CKPT_DIR, "ckpt-200"))
    got = read_tf_checkpoint(prefix)
    assert set(got) == set(ref)
    for k in ref:
        if k.startswith("save_counter"):
            assert int(np.asarray(got[k]).reshape(-1)[0]) == 3
            continue
        assert got[k].shape == ref[k].shape and got[k].dtype == ref[k].dtype
        np.testing.assert_array_equal(got[k], ref[k])  # unconstrained variables survive bit for bit
    assert "_CHECKPOINTABLE_OBJECT_GRAPH" in read_index(prefix)
    b = update_model_from_checkpoint(_model(), str(tmp_path / "log"))
    for pa, pb in zip(a.trainable_variables, b.trainable_variables):
        np.testing.assert_array_equal(pa.unconstrained, pb.unconstrained)
