"""Packed ensemble-MLP parameters with the reference's module surface.

Host-side mirror of ``mirl/torch_modules/mlp.py:63-178`` (``MLP``) over
``mirl/torch_modules/linear.py:114-165`` (``MyEnsembleLinear``).  The reference keeps E separate
``nn.Parameter`` per layer and ``torch.stack``s them on every forward (linear.py:146-149); here
all members of all layers live in ONE contiguous fp32 blob (layout: ``include/mirl_b200.h``,
``mb_mlp_desc``) that the CUDA kernels read directly and the fused Adam updates in place.
``state_dict`` / ``save`` / ``load`` speak the reference's key names
(``net.{2l}.weight_net{e}`` ``[in,out]``, ``net.{2l}.bias_net{e}`` ``[1,out]``,
``maxlogstd_net{e}`` / ``minlogstd_net{e}`` ``[1,D]``), so reference checkpoints load unchanged.
"""
import os.path as osp
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib


class PackedEnsembleMLP(nn.Module):
    def __init__(self, input_size, output_size, hidden_layers, ensemble_size, activation="relu",
                 output_activation="identity", module_name="mlp", probabilistic=False,
                 connection="simple", **linear_kwargs):
        super().__init__()
        if ensemble_size is None:
            raise NotImplementedError("mirl_b200 covers the ensemble path only (ensemble_size=None)")
        if connection != "simple":
            # densenet / resnet (linear.py:24-33) are not used by any config of the path
            raise NotImplementedError("connection=%r; only 'simple' is supported" % connection)
        if output_activation != "identity":
            raise NotImplementedError("output_activation=%r" % output_activation)
        if isinstance(activation, (list, tuple)):
            if len(set(activation)) != 1:
                raise NotImplementedError("per-layer activations")
            activation = activation[0]
        unknown = set(linear_kwargs) - {"with_bias", "init_bias_constant", "init_func_name",
                                        "init_kwargs"}
        if unknown or not linear_kwargs.get("with_bias", True):
            raise NotImplementedError("linear kwargs %s" % sorted(linear_kwargs))
        self.input_size, self.output_size = int(input_size), int(output_size)
        self.hidden_layers = [int(h) for h in hidden_layers]
        self.ensemble_size = int(ensemble_size)
        self.activation = activation
        self.module_name = module_name
        self.probabilistic = bool(probabilistic)
        self.layer_size = [self.input_size] + self.hidden_layers + [self.output_size]
        self.num_fc = len(self.layer_size) - 1
        self.desc = _lib.make_desc(self.layer_size, self.ensemble_size, activation)
        L = _lib.lib()
        self._w_off = [int(L.mb_mlp_weight_offset(self.desc, l)) for l in range(self.num_fc)]
        self._b_off = [int(L.mb_mlp_bias_offset(self.desc, l)) for l in range(self.num_fc)]
        self._ls_off = [int(L.mb_mlp_logstd_offset(self.desc, w)) for w in (0, 1)]
        n = int(L.mb_mlp_param_count(self.desc, int(self.probabilistic)))
        self.flat = nn.Parameter(torch.zeros(n, dtype=torch.float32))
        self._init_kwargs = dict(linear_kwargs)
        self.reset_parameters()

    # ---- views into the blob -----------------------------------------------------------------
    def weight(self, l, flat=None):
        """``[E, in, out]`` view of layer ``l`` (the reference's stacked ``weight_net*``)."""
        flat = self.flat.data if flat is None else flat
        E, i, o = self.ensemble_size, self.layer_size[l], self.layer_size[l + 1]
        return flat[self._w_off[l]:self._w_off[l] + E * i * o].view(E, i, o)

    def bias(self, l, flat=None):
        flat = self.flat.data if flat is None else flat
        E, o = self.ensemble_size, self.layer_size[l + 1]
        return flat[self._b_off[l]:self._b_off[l] + E * o].view(E, 1, o)

    def logstd_bound(self, which, flat=None):
        """``[E, 1, D]`` view of max (0) / min (1) log-std (pe_model.py:62-69)."""
        assert self.probabilistic
        flat = self.flat.data if flat is None else flat
        E, D = self.ensemble_size, self.output_size // 2
        return flat[self._ls_off[which]:self._ls_off[which] + E * D].view(E, 1, D)

    def member_slices(self, e0, e1):
        """Flat ``(start, stop)`` ranges holding members ``[e0, e1)`` (used to exchange
        member-sharded weights between ranks and for per-member snapshots)."""
        out = []
        E = self.ensemble_size
        for l in range(self.num_fc):
            wn = self.layer_size[l] * self.layer_size[l + 1]
            bn = self.layer_size[l + 1]
            out.append((self._w_off[l] + e0 * wn, self._w_off[l] + e1 * wn))
            out.append((self._b_off[l] + e0 * bn, self._b_off[l] + e1 * bn))
        if self.probabilistic:
            D = self.output_size // 2
            for w in (0, 1):
                out.append((self._ls_off[w] + e0 * D, self._ls_off[w] + e1 * D))
        assert 0 <= e0 <= e1 <= E
        return out

    def reset_parameters(self):
        """linear.py:69-88: ``orthogonal_`` on ``W.T`` per member, bias constant 0; log-std bounds
        +0.25 / -5 (pe_model.py:62-69)."""
        init = getattr(nn.init, self._init_kwargs.get("init_func_name", "orthogonal_"))
        kw = self._init_kwargs.get("init_kwargs", {})
        bias_const = self._init_kwargs.get("init_bias_constant", 0)
        with torch.no_grad():
            self.flat.zero_()
            for l in range(self.num_fc):
                i, o = self.layer_size[l], self.layer_size[l + 1]
                for e in range(self.ensemble_size):
                    w = torch.empty(i, o)
                    init(w.T, **kw)
                    self.weight(l)[e].copy_(w)
                    if bias_const is None:
                        nn.init.uniform_(self.bias(l)[e], -1 / i ** 0.5, 1 / i ** 0.5)
                    else:
                        self.bias(l)[e].fill_(bias_const)
            if self.probabilistic:
                self.logstd_bound(0).fill_(0.25)
                self.logstd_bound(1).fill_(-5.0)

    # ---- reference-compatible state dict --------------------------------------------------------
    def reference_state_dict(self, prefix="", flat=None):
        sd = OrderedDict()
        for e in range(self.ensemble_size):
            if self.probabilistic:
                sd["%smaxlogstd_net%d" % (prefix, e)] = self.logstd_bound(0, flat)[e]
                sd["%sminlogstd_net%d" % (prefix, e)] = self.logstd_bound(1, flat)[e]
        for l in range(self.num_fc):
            for e in range(self.ensemble_size):
                sd["%snet.%d.weight_net%d" % (prefix, 2 * l, e)] = self.weight(l, flat)[e]
                sd["%snet.%d.bias_net%d" % (prefix, 2 * l, e)] = self.bias(l, flat)[e]
        return sd

    def load_reference_state_dict(self, sd, prefix="", strict=True):
        own = self.reference_state_dict(prefix)
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k.startswith(prefix) and k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError("state dict mismatch: missing %s unexpected %s" % (missing, unexpected))
        with torch.no_grad():
            for k, view in own.items():
                if k in sd:
                    v = torch.as_tensor(sd[k])
                    if tuple(v.shape) != tuple(view.shape):
                        raise RuntimeError("shape mismatch for %s: %s vs %s"
                                           % (k, tuple(v.shape), tuple(view.shape)))
                    view.copy_(v)
        self.mark_weights_changed()
        return missing, unexpected

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        sd = self.reference_state_dict(prefix)
        if destination is not None:
            destination.update(sd)
            return destination
        return sd

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        """Accept the reference key layout (or our own single ``flat`` blob)."""
        if prefix + "flat" in state_dict:
            with torch.no_grad():
                self.flat.copy_(state_dict[prefix + "flat"])
            self.mark_weights_changed()
            return
        own = self.reference_state_dict(prefix)
        with torch.no_grad():
            for k, view in own.items():
                if k not in state_dict:
                    missing_keys.append(k)
                    continue
                v = torch.as_tensor(state_dict[k])
                if tuple(v.shape) != tuple(view.shape):
                    error_msgs.append("size mismatch for %s: %s vs %s"
                                      % (k, tuple(v.shape), tuple(view.shape)))
                    continue
                view.copy_(v)
        for k in state_dict:
            if k.startswith(prefix) and k not in own:
                unexpected_keys.append(k)
        self.mark_weights_changed()

    def mark_weights_changed(self):
        """Bump the epoch the derived device packs (tcgen05 operand pack) are keyed on.  Called
        by everything in this package that writes the blob outside torch's version counter."""
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1

    def weights_key(self):
        return (self.flat.data_ptr(), self.flat._version, getattr(self, "_weights_epoch", 0))

    # ---- mlp.py:111-168 snapshot / save / load (per-member files keyed by 'net{i}') ------------
    def get_snapshot(self, key_must_have=""):
        sd = self.reference_state_dict()
        if key_must_have == "":
            return sd
        return OrderedDict((k, v) for k, v in sd.items() if key_must_have in k)

    def load_snapshot(self, loaded_state_dict, key_must_have=""):
        if key_must_have != "":
            loaded_state_dict = {k: v for k, v in loaded_state_dict.items() if key_must_have in k}
        self.load_reference_state_dict(loaded_state_dict, "", strict=False)

    def _file(self, d, net_id):
        if net_id is None:
            return "", osp.join(d, "%s.pt" % self.module_name)
        name = "net%d" % net_id
        return name, osp.join(d, "%s_%s.pt" % (self.module_name, name))

    def save(self, save_dir, net_id=None):
        name, path = self._file(save_dir, net_id)
        torch.save({k: v.detach().cpu().clone() for k, v in self.get_snapshot(name).items()}, path)

    def load(self, load_dir, net_id=None):
        name, path = self._file(load_dir, net_id)
        if net_id is not None and not osp.exists(path):
            path = osp.join(load_dir, "%s.pt" % self.module_name)
        if not osp.exists(path):
            return False
        self.load_snapshot(torch.load(path, map_location="cpu"), name)
        return True

    def get_weight_decay(self, weight_decays=0):
        """mlp.py:170-178 / linear.py:161-165: sum_l wd_l * 0.5 * sum_e ||W_l^e||^2 (host-side
        helper for reporting; the training kernel folds the L2 gradient in)."""
        if not isinstance(weight_decays, (list, tuple)):
            weight_decays = [weight_decays] * self.num_fc
        assert len(weight_decays) == self.num_fc
        return sum((self.weight(l) ** 2).sum() * wd * 0.5 for l, wd in enumerate(weight_decays))

    def forward(self, x):
        raise RuntimeError("PackedEnsembleMLP has no eager forward; use the owning model/value "
                           "(CUDA kernels only, no PyTorch fallback)")
