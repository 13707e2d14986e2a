"""Drop-in mirrors of the reference's models package for the hot path (LNet, DNet, ENet).

``load_checkpoint`` / ``load_network`` / ``load_DNet`` keep the behaviour of the reference's
models/__init__.py:12-56 (strip ``module.``, drop ``low_res`` keys, strict=False; DNet from
``checkpoint['net_G_ema']``)."""
import torch


def _load(checkpoint_path):
    return torch.load(checkpoint_path, map_location=lambda storage, loc: storage)


def load_checkpoint(path, model):
    print("Load checkpoint from: {}".format(path))
    checkpoint = _load(path)
    try:
        s = checkpoint["state_dict"] if "arcface" not in path else checkpoint
        new_s = {}
        for k, v in s.items():
            if "low_res" in k:
                continue
            new_s[k.replace("module.", "")] = v
        model.load_state_dict(new_s, strict=False)
    except (KeyError, AttributeError, TypeError, RuntimeError):
        # the reference's fallback (models/__init__.py:24-26, a bare ``except``): the file holds a plain state_dict
        model.load_state_dict(checkpoint)
    return model


def load_network(args):
    """models/__init__.py:29-35 of the reference: LNet from ``args.LNet_path``, wrapped by ENet from ``args.ENet_path`` (whose
    ``low_res.*`` keys are skipped by load_checkpoint, so the LNet weights stay the ones just loaded); returns ``model.eval()``
    on the CPU like the reference - the caller moves it to the GPU (inference.py:250)."""
    from .ENet import ENet
    from .LNet import LNet
    L_net = LNet()
    L_net = load_checkpoint(args.LNet_path, L_net)
    E_net = ENet(lnet=L_net)
    model = load_checkpoint(args.ENet_path, E_net)
    return model.eval()


def load_DNet(args):
    from .DNet import DNet
    device = "cuda"
    net = DNet()
    checkpoint = torch.load(args.DNet_path, map_location=lambda storage, loc: storage)
    net.load_state_dict(checkpoint["net_G_ema"], strict=False)
    return net.to(device).eval()
