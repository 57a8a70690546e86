"""SAC on B200: mirrors SACImpl (d3rlpy/algos/torch/sac_impl.py:22-162).

The reference's CQLImpl *extends* SACImpl; here the relation is used the other way round: SAC is the CQL update graph
with zero importance-sampling groups (`n_action_samples = 0`: the critic input is the data rows only and the loss
kernels take their plain-TD branch), no alpha step, and the soft backup
`min_e Q'(s', a') - exp(log_temp) log pi(a'|s')` (sac_impl.py:148-162) always on."""
from __future__ import annotations

from .cql_impl import CQLImpl


class SACImpl(CQLImpl):
    def __init__(self, **kw):
        for k in ("alpha_learning_rate", "initial_alpha", "alpha_threshold", "conservative_weight", "n_action_samples",
                  "soft_q_backup"):
            kw.pop(k, None)
        super().__init__(alpha_learning_rate=0.0, n_action_samples=0, soft_q_backup=True, conservative_weight=0.0, **kw)

    def noise_layout(self, B):
        """Draw order of SAC._update (algos/sac.py:177-198): temperature step, target action, actor action."""
        A = self._action_size
        lay = {}
        if self._temp_learning_rate > 0:
            lay["temp"] = ("normal", (B, A))
        lay["soft"] = ("normal", (B, A))
        lay["actor"] = ("normal", (B, A))
        return lay

    def _checkpoint_views(self):
        views = super()._checkpoint_views()
        for k in ("_log_alpha", "_alpha_optim"):  # SACImpl has no alpha (torch_utility.get_state_dict walks attributes)
            views.pop(k, None)
        return views
