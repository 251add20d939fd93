"""Checkpoint compatibility with the reference: AFCM stores a network as the plain `state_dict()` of
the CPU module under `<save_dir>/<epoch>_net_<name>.pth` (models/base_model.py:144-161, names `G`, `G_ema`, `D`) and loads it
with `torch.load(..., map_location)` + `load_state_dict` (models/base_model.py:176-199).  The generator classes of this package
keep the reference's parameter and buffer names, so those files load unchanged; the functions below restate the file-name
convention, the `module.` prefix a DataParallel-wrapped save carries, and the `_metadata` handling."""
import os

import torch


def network_filename(epoch, name):
    """'%s_net_%s.pth' % (epoch, name)  (models/base_model.py:152, 185)."""
    return '%s_net_%s.pth' % (epoch, name)


def save_network(net, save_dir, epoch, name):
    """Writes the CPU state_dict of `net` in the reference's format and returns the path.  The module stays on its device
    (the reference moves it to the CPU and back; the file is the same)."""
    os.makedirs(save_dir, exist_ok=True)
    path = os.path.join(save_dir, network_filename(epoch, name))
    net = net.module if isinstance(net, torch.nn.DataParallel) else net
    torch.save({k: v.detach().to('cpu') for k, v in net.state_dict().items()}, path)
    return path


def load_network(net, path_or_dir, epoch=None, name=None, device=None, strict=True):
    """Loads a reference-format checkpoint into `net` (a generator of this package or of the reference).  `path_or_dir` is
    the file itself, or the directory together with `epoch` and `name`.  Keys saved from a DataParallel wrapper
    (`module.<key>`) are accepted."""
    path = path_or_dir if epoch is None else os.path.join(path_or_dir, network_filename(epoch, name))
    target = net.module if isinstance(net, torch.nn.DataParallel) else net
    if device is None:
        p = next(target.parameters(), None)
        device = p.device if p is not None else 'cpu'
    state = torch.load(path, map_location=str(device), weights_only=True)
    if hasattr(state, '_metadata'):
        del state._metadata
    if state and all(k.startswith('module.') for k in state):
        state = {k[len('module.'):]: v for k, v in state.items()}
    return target.load_state_dict(state, strict=strict)
