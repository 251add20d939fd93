"""CPU check of the warp-level recipe that packs filtered_lrelu sign codes out of mma.sync accumulator fragments into the
reference sign-tensor layout (tools/flr_sign_pack_model.py): for every phase shift the butterfly + funnel shift gives the
words the layout defines, and those bytes are the oracle's sign bytes for the same codes."""
import numpy as np
import pytest
import scipy.signal

from oracle import afcm_oracle as orc
from tools.flr_sign_pack_model import bytes_of, pack_chunk, pack_reference


@pytest.mark.parametrize('MB', [3, 5])
@pytest.mark.parametrize('sx', [0, 1, 2, 3])
def test_warp_packing_equals_definition(MB, sx):
    rng = np.random.RandomState(10 * MB + sx)
    code = rng.randint(0, 3, size=(16 * MB, 16)).astype(np.uint64)
    assert np.array_equal(pack_chunk(code, sx), pack_reference(code, sx))


def test_packed_bytes_are_the_oracle_sign_bytes():
    """Codes derived from the oracle's own pre-activation values (1 = negative, 2 = clamped), packed by the warp recipe, are
    the bytes of the oracle's sign tensor (first 16 rows, first 32 up-sampled columns of a plane; sx = 0 there)."""
    rng = np.random.RandomState(0)
    fu = scipy.signal.firwin(12, 0.4, width=0.3, fs=2).astype(np.float32)
    x = (rng.randn(1, 1, 38, 38) * 2).astype(np.float32)
    y, so, pre = orc.filtered_lrelu(x, fu, fu, None, up=2, down=2, padding=[9, 8, 9, 8], gain=np.sqrt(2), slope=0.2, clamp=1.0,
                                    write_signs=True, return_preact=True)
    pre = pre[0, 0]                                            # [up-sampled rows, columns], after gain, before slope / clamp
    act = np.where(pre < 0, pre * 0.2, pre)
    code = np.where(np.abs(act) > 1.0, 2, (pre < 0).astype(np.int64)).astype(np.uint64)
    words = pack_chunk(np.ascontiguousarray(code[:16, :48].T), 0)          # [J][V] for rows 0..15, columns 0..47
    assert np.array_equal(bytes_of(words), so[0, 0, :16, :8])
