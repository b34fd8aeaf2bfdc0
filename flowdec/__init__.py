"""Drop-in import names of the reference package (`flowdec.model.FlowModel`,
`flowdec.backbones.ncsnpp.NCSNpp`, ... as used by the reference's Hydra `_target_`s and
demo.ipynb), resolved to the B200-native implementation in `flowdec_b200`."""
import importlib
import sys

_ALIASES = {
    "flowdec.model": "flowdec_b200.model",
    "flowdec.backbones": "flowdec_b200.backbones",
    "flowdec.backbones.ncsnpp": "flowdec_b200.backbones.ncsnpp",
    "flowdec.data": "flowdec_b200.data",
    "flowdec.data.feature_extractors": "flowdec_b200.data.feature_extractors",
    "flowdec.data.sigma_models": "flowdec_b200.data.sigma_models",
    "flowdec.sampling": "flowdec_b200.sampling",
    "flowdec.sampling.solvers": "flowdec_b200.sampling.solvers",
    "flowdec.util": "flowdec_b200.util",
    "flowdec.util.other": "flowdec_b200.util.other",
}
for _alias, _target in _ALIASES.items():
    sys.modules[_alias] = importlib.import_module(_target)
from flowdec_b200 import model, backbones, data, sampling, util  # noqa: E402,F401
