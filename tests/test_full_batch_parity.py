"""GPU: elementwise parity of EVERY image of every BASELINE.json configuration (2-5) against the CPU oracle, through
the path the benchmark times (reslic_tcm_b200.pipeline.TcmEntropyPath: ctypes -> C ABI -> sm_100a kernels):

* per-slice launches (TCM's call pattern, src/models/reference/tcm.py:443-457 / :527-552) and whole-y launches;
* config 5 in NOISE mode with the explicit noise tensors (tcm.py:455 semantics: y_hat = y + u, likelihood at
  |y + u - mu|, ste_round output beside it) — the launch-size-dependent kernel variants (5 / 4 CTAs per SM, two-wave,
  two-tile balance) are picked by the full-size launches here, not by a small stand-in;
* the 8-GPU shard batches B = 3, 8, 2, 32, with the rate exchange attached (world 1) so that the
  collecting launch runs the instantiation a multi-GPU step runs.

Bit-exact: y_hat, z_hat, symbols, indexes.  Likelihoods: |L - L_ref| <= 1e-5 L_ref + 3e-7 (tests/util.py).  Per-image
bits: 1e-5 relative.  The oracle runs once per configuration (seconds on the host cores)."""
import pytest
import torch

from oracle import compressai_ref as cr
from reslic_tcm_b200 import dist as rdist
from reslic_tcm_b200 import synthetic
from reslic_tcm_b200.pipeline import TcmEntropyPath
from tests.util import assert_equal_exact, assert_lik_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_eb(params):
    eb = cr.EntropyBottleneckRef(synthetic.Z_CHANNELS)
    eb.matrices = [params[f"_matrix{i}"] for i in range(5)]
    eb.biases = [params[f"_bias{i}"] for i in range(5)]
    eb.factors = [params[f"_factor{i}"] for i in range(4)]
    eb.quantiles = params["quantiles"]
    return eb


def _path(params):
    p = TcmEntropyPath().to(DEV).eval()
    synthetic.load_eb_parameters(p.entropy_bottleneck, params)
    p.gaussian_conditional.scale_table = synthetic.scale_table(DEV)
    return p


def _compare(res, ref, c, images, what):
    """res: TcmEntropyPath result for `images` (a range into the oracle's batch); ref: oracle output of the whole batch."""
    sl = slice(images.start, images.stop)
    assert_equal_exact(res["y_hat"], ref["y_hat"][sl], f"{what}: y_hat (ste_round output)")
    assert_equal_exact(res["z_hat"], ref["z_hat"][sl], f"{what}: z_hat")
    if c.with_indexes:
        assert_equal_exact(res["symbols"], ref["symbols"][sl], f"{what}: symbols")
        assert_equal_exact(res["indexes"], ref["indexes"][sl], f"{what}: indexes")
    if c.training:
        assert_equal_exact(res["y_noisy"], ref["y_noisy"][sl], f"{what}: quantize('noise') output")
    assert_lik_close(res["likelihoods"]["y"], ref["y_lik"][sl], what=f"{what}: y likelihood")
    assert_lik_close(res["likelihoods"]["z"], ref["z_lik"][sl], what=f"{what}: z likelihood")
    bits_ref = (cr.per_image_bits(ref["y_lik"][sl]) + cr.per_image_bits(ref["z_lik"][sl]))
    assert torch.allclose(res["bits"].cpu(), bits_ref, rtol=1e-5, atol=0), f"{what}: per-image bits"
    bpp = float(res["bits"].sum()) / (len(images) * c.num_pixels_per_image)
    bpp_ref = float(bits_ref.sum()) / (len(images) * c.num_pixels_per_image)
    assert abs(bpp - bpp_ref) <= 1e-5 * bpp_ref, f"{what}: bpp {bpp} vs {bpp_ref}"


@pytest.mark.parametrize("cfg", [2, 3, 4, 5])
def test_every_image_of_the_config_matches_the_oracle(cfg):
    c = synthetic.CONFIGS[cfg]
    params = synthetic.eb_parameters()
    batch = synthetic.make_batch(cfg, range(c.batch), with_noise=c.training)
    with torch.no_grad():
        ref = cr.tcm_entropy_step(batch["y"], batch["mu"], batch["sigma"], batch["z"], _ref_eb(params), synthetic.scale_table(),
                                  training=c.training, with_indexes=c.with_indexes,
                                  noise_y=batch.get("noise_y"), noise_z=batch.get("noise_z"))
    if c.training:
        ref["y_noisy"] = batch["y"] + batch["noise_y"]             # EntropyModel.quantize(y, "noise"): means ignored (App. A.1)
    kw = dict(training=c.training, with_indexes=c.with_indexes, num_pixels=c.num_pixels_per_image)

    def run(images, **over):
        path = _path(params)
        dev = {k: batch[k][images.start:images.stop].to(DEV) for k in batch}
        res = path.forward(dev["y"], dev["mu"], dev["sigma"], dev["z"], noise_y=dev.get("noise_y"), noise_z=dev.get("noise_z"),
                           **dict(kw, **over))
        torch.cuda.synchronize()
        return res

    whole = range(c.batch)
    _compare(run(whole), ref, c, whole, f"config {cfg}, {c.batch} images, per-slice launches")
    _compare(run(whole, fuse_slices=True), ref, c, whole, f"config {cfg}, {c.batch} images, whole-y launch")
    # the 8-GPU shards (B = 3 / 8 / 2 / 32): first, a middle and the last rank; the collecting launch also publishes
    for rank in (0, 3, 7):
        images = rdist.shard_range(c.batch, rank, 8)
        ex = rdist.PeerRateExchange(torch.device(DEV), ring=4)
        ex.set_static(pixels=len(images) * c.num_pixels_per_image, images=len(images))
        res = run(images, exchange=ex)
        _compare(res, ref, c, images, f"config {cfg}, rank {rank} of 8 ({len(images)} images), per-slice launches")
        row = ex.read(1)[0].tolist()
        ex.check()
        assert row[0] == float(res["bits"].double().sum()) and row[3] == len(images)
    assert len(rdist.shard_range(c.batch, 3, 8)) == {2: 3, 3: 8, 4: 2, 5: 32}[cfg]
