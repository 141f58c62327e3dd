"""cobaya ``Likelihood`` plugin backed by the B200 path (drop-in for
victor/likelihoods/CCFLikelihood.py:6-42).

With cobaya installed this subclasses ``cobaya.likelihood.Likelihood``; without it (this image)
a small stand-in base class provides the part of the protocol the plugin uses -- class
attributes from the yaml defaults next to this file, keyword overrides, then ``initialize()``
-- so the plugin can be driven by any sampler loop that calls ``calculate``.

One instance owns one GPU context.  Under ``mpirun -n N cobaya-run`` (one chain per rank) set
``device`` per rank, or leave it unset to use ``LOCAL_RANK`` / ``OMPI_COMM_WORLD_LOCAL_RANK``
modulo the number of visible GPUs.
"""
import os

import yaml

try:  # pragma: no cover - cobaya is not part of this image
    from cobaya.likelihood import Likelihood
except ImportError:
    class Likelihood:
        """Minimal stand-in for cobaya.likelihood.Likelihood."""

        def __init__(self, info=None, **kwargs):
            defaults_fn = os.path.splitext(os.path.abspath(__file__))[0] + ".yaml"
            with open(defaults_fn) as fh:
                attrs = yaml.full_load(fh) or {}
            attrs.update(info or {})
            attrs.update(kwargs)
            for key, value in attrs.items():
                setattr(self, key, value)
            self.initialize()

        def initialize(self):
            pass

from victor_b200 import CCFFit


def _rank_device():
    for var in ("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", "MV2_COMM_WORLD_LOCAL_RANK"):
        if var in os.environ:
            from victor_b200 import _lib
            ndev = max(1, _lib.load().vb200_device_count())
            return int(os.environ[var]) % ndev
    return None


class CCFLikelihood(Likelihood):

    def initialize(self):
        """Build the CCFFit object (and, lazily, its GPU context)."""
        if self.model is None or self.data is None:
            # read the blocks from the configuration file, as the reference does (:12-21)
            if os.path.isfile(self.config_file):
                with open(self.config_file) as fh:
                    info = yaml.full_load(fh)
                self.model = info["model"]
                self.data = info["data"]
            else:
                raise KeyError(f"config file {self.config_file} not found")
        device = getattr(self, "device", None)
        if device is None:
            device = _rank_device()
        self.ccf = CCFFit(self.model, self.data, device=device)

    def get_can_provide_params(self):
        return ["fsigma8"]

    def calculate(self, state, want_derived=True, **params_values):
        """Fill ``state['logp']`` and the derived chi-square (reference :32-42)."""
        lnlike, chisq = self.ccf.log_likelihood(params_values)
        state["logp"] = lnlike
        state["derived"] = {"chi2_ccf_correct": chisq}
