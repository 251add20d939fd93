"""conv2d_resample -- 2-D convolution with optional up / down-sampling, the operator the discriminator and the CoModGAN baseline
generator are built on (reference: models/networks/CoModGAN/torch_utils/ops/conv2d_resample.py:59-156, called from
CM/layers.py:157 and CM/layers.py:20-77).  Same signature and the same results; which of blur / strided convolution / transposed
convolution runs, in which order and with which padding, follows the reference's case analysis (cited per branch), expressed
here as a short plan over two closures.  The pieces are this package's native operators."""
import torch

from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _get_filter_size, _parse_padding


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """(:27-55) `flip_weight=False` asks for a true convolution, i.e. the kernel is flipped before the correlation."""
    if not flip_weight and (w.shape[2] > 1 or w.shape[3] > 1):
        w = w.flip([2, 3])
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def _resample_padding(padding, fw, fh, up, down):
    """Padding relative to the up / down-sampled image (:94-104): the filter's support is split around the sampling grid."""
    px0, px1, py0, py1 = _parse_padding(padding)
    for factor, lo_extra, hi_extra in ((up, up - 1, -up), (down, 1 - down, -down)):
        if factor > 1:
            px0 += (fw + lo_extra) // 2
            px1 += (fw + hi_extra) // 2
            py0 += (fh + lo_extra) // 2
            py1 += (fh + hi_extra) // 2
    return px0, px1, py0, py1


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1 and isinstance(groups, int) and groups >= 1
    out_channels, in_per_group, kh, kw = (int(v) for v in w.shape)
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _resample_padding(padding, fw, fh, up, down)
    pointwise = kh == 1 and kw == 1

    def blur(t, **kwargs):                                   # upfirdn2d with this call's filter
        return upfirdn2d.upfirdn2d(x=t, f=f, flip_filter=flip_filter, **kwargs)

    def conv(t, **kwargs):                                   # the convolution proper
        return _conv2d_wrapper(x=t, w=w, groups=groups, flip_weight=flip_weight, **kwargs)

    if up == 1 and down > 1:
        if pointwise:                                        # (:107-110) a 1x1 kernel commutes with decimation: decimate first
            return conv(blur(x, down=down, padding=[px0, px1, py0, py1]))
        return conv(blur(x, padding=[px0, px1, py0, py1]), stride=down)            # (:119-122) blur, then strided convolution
    if up > 1 and down == 1 and pointwise:                   # (:113-116) ... and with up-sampling: convolve first
        return blur(conv(x), up=up, padding=[px0, px1, py0, py1], gain=up ** 2)
    if up > 1:
        # (:125-141) transposed strided convolution, then the blur (and the decimation, if any)
        wt = w.transpose(0, 1) if groups == 1 else \
            w.reshape(groups, out_channels // groups, in_per_group, kh, kw).transpose(1, 2).reshape(groups * in_per_group,
                                                                                                     out_channels // groups, kh, kw)
        qx0, qx1, qy0, qy1 = px0 - (kw - 1), px1 - (kw - up), py0 - (kh - 1), py1 - (kh - up)
        tx, ty = max(min(-qx0, -qx1), 0), max(min(-qy0, -qy1), 0)
        y = _conv2d_wrapper(x=x, w=wt, stride=up, padding=[ty, tx], groups=groups, transpose=True, flip_weight=(not flip_weight))
        y = blur(y, padding=[qx0 + tx, qx1 + tx, qy0 + ty, qy1 + ty], gain=up ** 2)
        return blur(y, down=down) if down > 1 else y
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:  # (:144-146) no resampling: the convolution pads by itself
        return conv(x, padding=[py0, px0])
    # (:149-153) generic composition: explicit padding, convolution (up == down == 1 here)
    return conv(upfirdn2d.upfirdn2d(x=x, f=None, padding=[px0, px1, py0, py1], flip_filter=flip_filter))
