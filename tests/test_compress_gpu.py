"""GPU: compress()/decompress() of the drop-in modules — symbols and indexes from the fused
kernels, bitstream from the host rANS coder — as TCM.compress / TCM.decompress drive them
(src/models/reference/tcm.py:502-568, 590-635)."""
import pytest
import torch

from oracle import compressai_ref as cr
from reslic_tcm_b200 import EntropyBottleneck, GaussianConditional, rans, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_gaussian_conditional_compress_decompress_round_trip():
    gc = GaussianConditional(None).to(DEV).eval()
    assert gc.update_scale_table(cr.get_scale_table()) is True
    batch = synthetic.make_batch(1, range(2), y_hw=(8, 8), z_hw=(2, 2))
    y, mu, sg = (batch[k][:, :64].to(DEV) for k in ("y", "mu", "sigma"))
    with torch.no_grad():
        idx = gc.build_indexes(sg)
        strings = gc.compress(y, idx, means=mu)
        y_hat = gc.decompress(strings, idx, means=mu)
        ref = gc.quantize(y, "dequantize", mu)
    assert len(strings) == 2 and all(isinstance(s, bytes) for s in strings)
    assert torch.equal(y_hat, ref)
    with pytest.raises(ValueError):
        gc.compress(y, idx[:, :32], means=mu)
    with pytest.raises(ValueError):
        GaussianConditional(None).to(DEV).compress(y, idx, means=mu)      # CDFs not initialised


def test_entropy_bottleneck_compress_decompress_round_trip():
    torch.manual_seed(3)
    eb = EntropyBottleneck(16).to(DEV).eval()
    with torch.no_grad():
        eb.quantiles[:, 0, 1] = torch.randn(16, device=DEV)
        eb.quantiles[:, 0, 0] = eb.quantiles[:, 0, 1] - 9.0
        eb.quantiles[:, 0, 2] = eb.quantiles[:, 0, 1] + 9.0
    assert eb.update() is True
    z = 3.0 * torch.randn(3, 16, 4, 5, device=DEV)
    z[0, 0, 0, 0] = 45.0                                                    # outside the table: bypass escape
    with torch.no_grad():
        strings = eb.compress(z)
        z_hat = eb.decompress(strings, z.size()[-2:])
        med = eb._get_medians().reshape(1, -1, 1, 1)
        ref = torch.round(z - med) + med
    assert z_hat.shape == z.shape and torch.equal(z_hat, ref)


def test_tcm_style_sliced_stream():
    """One BufferedRansEncoder over the 5 slices, one RansDecoder reading them back in order,
    with GPU int32 tensors instead of Python lists (tcm.py:527-565, 604-623)."""
    gc = GaussianConditional(None).to(DEV).eval()
    gc.update_scale_table(cr.get_scale_table())
    batch = synthetic.make_batch(2, range(1), y_hw=(8, 8), z_hw=(2, 2))
    y, mu, sg = (batch[k].to(DEV) for k in ("y", "mu", "sigma"))
    cdf, lengths, offsets = gc.quantized_cdf, gc.cdf_length, gc.offset
    enc = rans.BufferedRansEncoder()
    syms, idxs = [], []
    with torch.no_grad():
        for k in range(5):
            sl = slice(64 * k, 64 * (k + 1))
            r = gc.forward_fused(y[:, sl], sg[:, sl], mu[:, sl], want=("sym", "idx"))
            syms.append(r.sym); idxs.append(r.idx)
            enc.encode_with_indexes(r.sym, r.idx, cdf, lengths, offsets)
    stream = enc.flush()
    dec = rans.RansDecoder()
    dec.set_stream(stream)
    with torch.no_grad():
        for k in range(5):
            sl = slice(64 * k, 64 * (k + 1))
            idx = gc.build_indexes(sg[:, sl])
            assert torch.equal(idx, idxs[k])
            rv = dec.decode_stream_tensor(idx, cdf, lengths, offsets).reshape(idx.shape).to(DEV)
            assert torch.equal(rv, syms[k])
            y_hat = gc.dequantize(rv, mu[:, sl])
            assert torch.equal(y_hat, gc.quantize(y[:, sl], "dequantize", mu[:, sl]))
    bits_model = float(gc.forward_fused(y, sg, mu, want=("bits",)).bits.sum())
    assert len(stream) * 8 < 1.05 * bits_model + 256


def test_device_side_slots_match_the_host_lookup_and_give_identical_strings():
    """reslic_rans_slots_u32 against its host restatement (bit-exact slots, same escape set), the strings of
    GaussianConditional.compress against the host-lookup path, and the fallback when the escape list overflows."""
    from reslic_tcm_b200 import ops, rans
    from tests.util import rans_slots_reference

    gc = GaussianConditional(None).to(DEV)
    gc.update_scale_table(cr.get_scale_table())
    gen = torch.Generator().manual_seed(21)
    shape = (3, 16, 12, 10)
    sigma = torch.exp(torch.empty(shape).uniform_(-3, 4, generator=gen))
    mu = torch.randn(shape, generator=gen)
    y = mu + sigma * torch.randn(shape, generator=gen)
    y.view(-1)[::501] += 5000.0                       # escapes
    y.view(-1)[7::733] -= 9000.0
    yd, md, sd = y.to(DEV), mu.to(DEV), sigma.to(DEV)
    idx = gc.build_indexes(sd)
    sym = gc.quantize(yd, "symbols", md)
    slots, esc_pos, esc_raw, status = ops.rans_slots(sym, idx, gc._quantized_cdf, gc._cdf_length, gc._offset)
    ref_slots, ref_pos, ref_raw = rans_slots_reference(sym.cpu(), idx.cpu(), gc._quantized_cdf.cpu(), gc._cdf_length.cpu(),
                                                       gc._offset.cpu())
    assert torch.equal(slots.cpu(), ref_slots)
    n_esc = int(status[0])
    assert int(status[1]) == 0 and n_esc == ref_pos.numel() > 10
    pos, order = torch.sort(esc_pos[:n_esc].cpu())
    assert torch.equal(pos, ref_pos) and torch.equal(esc_raw[:n_esc].cpu()[order], ref_raw)
    strings = gc.compress(yd, idx, md)                 # device-side lookup
    host = rans.encode_with_indexes_batch(sym, idx, gc._quantized_cdf, gc._cdf_length, gc._offset)
    assert strings == host
    back = gc.decompress(strings, idx, means=md)
    assert torch.equal(back, gc.quantize(yd, "dequantize", md))
    # overflowing escape list -> None (compress falls back to the host lookup)
    small = ops.rans_slots(sym, idx, gc._quantized_cdf, gc._cdf_length, gc._offset, esc_capacity=4)
    assert rans.encode_slots_batch(*small) is None
    # an index outside the tables is reported, not read
    bad = idx.clone()
    bad.view(-1)[3] = 64
    with pytest.raises(ValueError, match="index out of range"):
        rans.encode_slots_batch(*ops.rans_slots(sym, bad, gc._quantized_cdf, gc._cdf_length, gc._offset))
