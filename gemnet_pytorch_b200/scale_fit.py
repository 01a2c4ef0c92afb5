"""Scale factors and their variance fit (reference gemnet/model/layers/scaling.py:7-174, used by
fit_scaling.py:93-160).

`ScalingFactor` is the module the blocks own: a scalar non-trainable parameter read from `scale_file` (JSON, key = the
factor's name).  In the kernels the value is folded into epilogues as a float argument (row H13, `value()`); the module
call `ScalingFactor(x_ref, y)` keeps the reference signature for layer-level use.

`AutomaticFit` / `AutoScaleFit` restate the reference's fitting protocol so that the unmodified `fit_scaling.py` loop
works on top of this package when `gemnet.model.layers.scaling` is aliased to this module: factors are
fitted ONE AT A TIME in creation order — `set2fitmode()` before the model is built, then for the active factor run
batches through `GemNet.forward` (which observes Var(x_ref) and Var(y) at that factor, on the GPU), call
`AutomaticFit.activeVar.fit()` (factor *= sqrt(Var_in / Var_out), written to the JSON file), which activates the next one,
until `fitting_completed()`.  While a factor is being fitted `GemNet.forward` takes the layer-by-layer kernel path
(`forward_autograd`): the chain-fused kernels keep the intermediate tensors on chip, where there is nothing to observe.
"""
import json
import logging

import numpy as np
import torch


def _check_json_path(path):
    if not path.endswith(".json"):
        raise UserWarning(f"Path {path} is not a json-path.")


def read_json(path):
    _check_json_path(path)
    with open(path, "r") as f:
        return json.load(f)


def write_json(path, data):
    _check_json_path(path)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(data, f, ensure_ascii=False, indent=4)


def update_json(path, data):
    content = read_json(path)
    content.update(data)
    write_json(path, content)


def read_value_json(path, key):
    return read_json(path).get(key)


class AutomaticFit:
    """One fitted scalar.  Class-level state (as in the reference, scaling.py:12-14): the factor being fitted
    (`activeVar`), the ones waiting (`queue`, `None` once everything is fitted) and whether newly created factors
    should enrol (`fitting_mode`)."""

    activeVar = None
    queue = None
    fitting_mode = False

    def __init__(self, variable, scale_file, name):
        self.variable, self.scale_file, self._name = variable, scale_file, name
        self._fitted = False
        self.load_maybe()
        if AutomaticFit.fitting_mode and not self._fitted:
            if AutomaticFit.activeVar is None:      # the first unfitted factor becomes active, the rest queue up
                AutomaticFit.activeVar = self
                AutomaticFit.queue = []
            else:
                self._add2queue()

    # called on the class: AutomaticFit.set2fitmode() / .reset() / .fitting_completed()  (fit_scaling.py:93,153)
    @staticmethod
    def reset():
        AutomaticFit.activeVar = None
        AutomaticFit.all_processed = False

    @staticmethod
    def fitting_completed():
        return AutomaticFit.queue is None

    @staticmethod
    def set2fitmode():
        AutomaticFit.reset()
        AutomaticFit.fitting_mode = True

    @staticmethod
    def leave_fitmode():
        """not in the reference: back to the normal mode (tests build several models in one process)"""
        AutomaticFit.reset()
        AutomaticFit.queue = None
        AutomaticFit.fitting_mode = False

    def _add2queue(self):
        if any(v._name == self._name for v in AutomaticFit.queue):
            raise ValueError(f"Variable with the same name ({self._name}) was already added to queue!")
        AutomaticFit.queue.append(self)

    def set_next_active(self):
        if not AutomaticFit.queue:
            AutomaticFit.queue = None
            AutomaticFit.activeVar = None
            return
        AutomaticFit.activeVar = AutomaticFit.queue.pop(0)

    def load_maybe(self):
        """take the value from the file when it has one (then this factor counts as fitted)"""
        if self.scale_file is None:
            return
        value = read_value_json(self.scale_file, self._name)
        if value is None:
            logging.info(f"Initialize variable {self._name} to {float(self.variable):.3f}")
            return
        self._fitted = True
        with torch.no_grad():
            self.variable.copy_(torch.tensor(value))


class AutoScaleFit(AutomaticFit):
    """Variance-preserving fit: the factor alpha with Var(alpha * y) = Var(x_ref), variances = mean over the feature
    columns of the per-column variance over the rows, sample-weighted over the observed batches (scaling.py:84-147)."""

    def __init__(self, variable, scale_file, name):
        super().__init__(variable, scale_file, name)
        if not self._fitted:
            self._init_stats()

    def _init_stats(self):
        self.variance_in = 0.0
        self.variance_out = 0.0
        self.nSamples = 0

    def is_active(self):
        return AutomaticFit.activeVar is self and not self._fitted

    def observe(self, x, y):
        if not self.is_active():
            return
        n = y.shape[0]
        self.variance_in += float(torch.var(x.detach().float(), dim=0).mean()) * n
        self.variance_out += float(torch.var(y.detach().float(), dim=0).mean()) * n
        self.nSamples += n

    def fit(self):
        if AutomaticFit.activeVar is not self:
            return
        if self.variance_in == 0:
            raise ValueError(f"Did not track the variable {self._name}. Add observe calls to track the variance "
                             "before and after.")
        var_in, var_out = self.variance_in / self.nSamples, self.variance_out / self.nSamples
        ratio = var_out / var_in
        value = np.sqrt(1 / ratio, dtype="float32")
        logging.info(f"Variable: {self._name}, Var_in: {var_in:.3f}, Var_out: {var_out:.3f}, Ratio: {ratio:.3f} "
                     f"=> Scaling factor: {value:.3f}")
        with torch.no_grad():
            self.variable.copy_(self.variable * float(value))
        self._fitted = True
        if self.scale_file is not None:
            update_json(self.scale_file, {self._name: float(self.variable)})
        self.set_next_active()


class ScalingFactor(torch.nn.Module):
    """y * scale_factor (scaling.py:150-174).  `value()` is the Python float the kernels fold into their epilogues;
    `observe(x_ref, y)` is the hook the blocks call after a kernel has produced the scaled y."""

    def __init__(self, scale_file, name, device=None):
        super().__init__()
        self.scale_factor = torch.nn.Parameter(torch.tensor(1.0, device=device), requires_grad=False)
        self._name = name
        if scale_file is not None:
            _check_json_path(scale_file)
        self.autofit = AutoScaleFit(self.scale_factor, scale_file, name)
        self._cached = None

    def value(self):
        """Python float of the factor; cached per parameter version (one device->host read)."""
        v = self.scale_factor._version
        if self._cached is None or self._cached[0] != v or self._cached[2] != self.scale_factor.data_ptr():
            self._cached = (v, float(self.scale_factor.detach().cpu()), self.scale_factor.data_ptr())
        return self._cached[1]

    def observe(self, x_ref, y):
        """x_ref: tensor, or a callable returning it (evaluated only while THIS factor is being fitted)"""
        if self.autofit.is_active():
            self.autofit.observe(x_ref() if callable(x_ref) else x_ref, y)

    def forward(self, x_ref, y):
        y = y * self.scale_factor
        self.observe(x_ref, y)
        return y


def fitting_active():
    """is some factor waiting for observations right now (GemNet.forward then takes the observable layer path)"""
    return AutomaticFit.fitting_mode and AutomaticFit.activeVar is not None
