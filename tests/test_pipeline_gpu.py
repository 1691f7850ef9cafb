"""GPU: the whole per-batch entropy pass (1 bottleneck + 5 slice launches, rate accumulated by
the kernels) against the oracle's tcm_entropy_step, in per-slice, fused-slice and CUDA-graph
form, eval and training mode."""
import pytest
import torch

from oracle import compressai_ref as cr
from reslic_tcm_b200 import ops, synthetic
from reslic_tcm_b200.pipeline import HostPipeline, TcmEntropyPath
from tests.util import assert_equal_exact, assert_lik_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(cfg, B, y_hw, z_hw, with_noise=False):
    batch = synthetic.make_batch(cfg, range(B), y_hw=y_hw, z_hw=z_hw, with_noise=with_noise)
    params = synthetic.eb_parameters()
    path = TcmEntropyPath().to(DEV).eval()
    synthetic.load_eb_parameters(path.entropy_bottleneck, params)
    path.gaussian_conditional.scale_table = synthetic.scale_table(DEV)
    ref_eb = cr.EntropyBottleneckRef(192)
    ref_eb.matrices = [params[f"_matrix{i}"] for i in range(5)]
    ref_eb.biases = [params[f"_bias{i}"] for i in range(5)]
    ref_eb.factors = [params[f"_factor{i}"] for i in range(4)]
    ref_eb.quantiles = params["quantiles"]
    return batch, path, ref_eb


def _compare(res, ref, num_pixels_per_image, B, with_indexes):
    assert_equal_exact(res["y_hat"], ref["y_hat"], "y_hat")
    assert_equal_exact(res["z_hat"], ref["z_hat"], "z_hat")
    if with_indexes:
        assert_equal_exact(res["symbols"], ref["symbols"], "symbols")
        assert_equal_exact(res["indexes"], ref["indexes"], "indexes")
    assert_lik_close(res["likelihoods"]["y"], ref["y_lik"], what="y likelihood")
    assert_lik_close(res["likelihoods"]["z"], ref["z_lik"], what="z likelihood")
    bits_ref = cr.per_image_bits(ref["y_lik"]) + cr.per_image_bits(ref["z_lik"])
    assert torch.allclose(res["bits"].cpu(), bits_ref, rtol=1e-5)
    bpp = float(res["bits"].sum()) / (num_pixels_per_image * B)
    assert bpp == pytest.approx(float(ref["bpp"]), rel=1e-5)      # loss.py:24-27


@pytest.mark.parametrize("fuse", [False, True])
def test_eval_pass_matches_oracle_step(fuse):
    B, y_hw, z_hw = 3, (16, 16), (4, 4)
    batch, path, ref_eb = _setup(1, B, y_hw, z_hw)
    dev = {k: v.to(DEV) for k, v in batch.items()}
    res = path(dev["y"], dev["mu"], dev["sigma"], dev["z"], with_indexes=True, num_pixels=256 * 256, fuse_slices=fuse)
    ref = cr.tcm_entropy_step(batch["y"], batch["mu"], batch["sigma"], batch["z"], ref_eb, synthetic.scale_table(),
                              num_pixels=256 * 256 * B)
    _compare(res, ref, 256 * 256, B, True)


def test_training_pass_with_explicit_noise_matches_oracle_step():
    B, y_hw, z_hw = 2, (8, 8), (2, 2)
    batch, path, ref_eb = _setup(5, B, y_hw, z_hw, with_noise=True)
    dev = {k: v.to(DEV) for k, v in batch.items()}
    res = path(dev["y"], dev["mu"], dev["sigma"], dev["z"], training=True, noise_y=dev["noise_y"],
               noise_z=dev["noise_z"], num_pixels=128 * 128)
    ref = cr.tcm_entropy_step(batch["y"], batch["mu"], batch["sigma"], batch["z"], ref_eb, synthetic.scale_table(),
                              training=True, with_indexes=False, num_pixels=128 * 128 * B,
                              noise_y=batch["noise_y"], noise_z=batch["noise_z"])
    _compare(res, ref, 128 * 128, B, False)
    assert_equal_exact(res["y_noisy"], batch["y"] + batch["noise_y"], "noisy y")


def test_cuda_graph_replay_tracks_new_inputs():
    B, y_hw, z_hw = 2, (16, 16), (4, 4)
    batch, path, ref_eb = _setup(1, B, y_hw, z_hw)
    dev = {k: v.to(DEV).clone() for k, v in batch.items()}
    graph, res = path.capture(dev["y"], dev["mu"], dev["sigma"], dev["z"], with_indexes=True, num_pixels=256 * 256)
    other = synthetic.make_batch(2, range(B), y_hw=y_hw, z_hw=z_hw)
    for k in ("y", "mu", "sigma", "z"):
        dev[k].copy_(other[k])
    for _ in range(3):                       # replays must not accumulate stale rate
        graph.replay()
    torch.cuda.synchronize()
    ref = cr.tcm_entropy_step(other["y"], other["mu"], other["sigma"], other["z"], ref_eb, synthetic.scale_table(),
                              num_pixels=256 * 256 * B)
    _compare(res, ref, 256 * 256, B, True)


def test_bits_accumulate_flag():
    g = torch.Generator().manual_seed(4)
    y = torch.randn(3, 8, 4, 4, generator=g).to(DEV)
    s = (torch.rand(3, 8, 4, 4, generator=g) + 0.2).to(DEV)
    bits = torch.full((3,), 100.0, dtype=torch.float64, device=DEV)
    one = ops.gc_forward(y, s, None, want=("bits",)).bits.clone()
    ops.gc_forward(y, s, None, want=("bits",), out={"bits": bits, "bits_accumulate": True})
    assert torch.allclose(bits, one + 100.0, rtol=1e-12)
    ops.gc_forward(y, s, None, want=("bits",), out={"bits": bits})
    assert torch.equal(bits, one)


def test_host_pipeline_streams_batches_and_matches_the_device_pass():
    """HostPipeline: pinned host latents in, host symbols/indexes/bits out, chunked over images and
    double-buffered over consecutive batches.  Four different batches are enqueued back to back without
    a synchronisation in between (so uploads, kernels and downloads of neighbouring batches overlap and
    every buffer slot is reused); each must equal the device-resident pass on the same data."""
    B, y_hw, z_hw = 5, (16, 8), (4, 2)
    _, path, _ = _setup(2, B, y_hw, z_hw)
    hp = HostPipeline(path, B, y_hw, z_hw, with_indexes=True, chunks=3, device=DEV)
    batches = [synthetic.make_batch(2, range(10 * k, 10 * k + B), y_hw=y_hw, z_hw=z_hw, pin=True) for k in range(4)]
    got = []
    for k, host in enumerate(batches):
        got.append(hp.run(host))
        # results of a slot stay valid until `depth` further calls: copy batch k-1 before it is overwritten
        if k >= 1:
            got[k - 1]["done"].synchronize()
            got[k - 1] = {n: got[k - 1][n].clone() for n in ("bits", "symbols", "indexes")}
    hp.synchronize()
    got[-1] = {n: got[-1][n].clone() for n in ("bits", "symbols", "indexes")}
    for k, host in enumerate(batches):
        ref = path.forward(host["y"].to(DEV), host["mu"].to(DEV), host["sigma"].to(DEV), host["z"].to(DEV),
                           with_indexes=True)
        assert torch.equal(got[k]["symbols"], ref["symbols"].cpu()), k
        assert torch.equal(got[k]["indexes"], ref["indexes"].cpu()), k
        assert torch.allclose(got[k]["bits"], ref["bits"].cpu(), rtol=1e-6), k      # chunking regroups fp32 partials


def test_host_pipeline_training_mode_uses_the_callers_philox_seed():
    """Noise mode through HostPipeline: with one chunk the launches are those of the device-resident pass, so the
    same seed must give the same per-image bits; another seed must not."""
    B, y_hw, z_hw = 4, (16, 8), (4, 2)
    _, path, _ = _setup(5, B, y_hw, z_hw)
    host = synthetic.make_batch(5, range(B), y_hw=y_hw, z_hw=z_hw, pin=True)
    ref = path.forward(host["y"].to(DEV), host["mu"].to(DEV), host["sigma"].to(DEV), host["z"].to(DEV),
                       training=True, seed=77)["bits"].clone().cpu()
    hp = HostPipeline(path, B, y_hw, z_hw, with_indexes=False, training=True, chunks=1, device=DEV, seed=77)
    out = hp.run(host)
    out["done"].synchronize()
    assert torch.allclose(out["bits"], ref, rtol=1e-6)
    hp2 = HostPipeline(path, B, y_hw, z_hw, with_indexes=False, training=True, chunks=1, device=DEV, seed=78)
    out2 = hp2.run(host)
    out2["done"].synchronize()
    assert not torch.allclose(out2["bits"], ref, rtol=1e-6)


def test_host_pipeline_packed_slots_give_the_strings_of_the_symbol_index_path():
    """packed_slots: the rANS lookup runs on the device and one 32-bit slot per symbol is downloaded; the strings coded
    from them must be byte-identical to those coded from the downloaded symbols and indexes."""
    from reslic_tcm_b200 import rans

    B, y_hw, z_hw = 5, (16, 8), (4, 2)
    _, path, _ = _setup(2, B, y_hw, z_hw)
    gc = path.gaussian_conditional
    gc.update()
    host = synthetic.make_batch(2, range(B), y_hw=y_hw, z_hw=z_hw, pin=True)
    host["y"].view(-1)[::4099] += 30000.0                                  # a few escapes, spread over the chunks
    plain = HostPipeline(path, B, y_hw, z_hw, with_indexes=True, chunks=2, device=DEV)
    packed = HostPipeline(path, B, y_hw, z_hw, with_indexes=True, chunks=2, device=DEV, packed_slots=True)
    assert packed.d2h_bytes < 0.6 * plain.d2h_bytes
    a = plain.run(host)
    b = packed.run(host)
    a["done"].synchronize()
    b["done"].synchronize()
    want = rans.encode_with_indexes_batch(a["symbols"], a["indexes"], gc._quantized_cdf, gc._cdf_length, gc._offset)
    got = packed.strings(b)
    assert int(b["slot_status"][:, 0].sum()) > 0 and got == want
    assert torch.equal(a["bits"], b["bits"])
