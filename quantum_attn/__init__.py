"""Drop-in alias: ``import quantum_attn`` gives the B200-native package under the reference's module names
(``quantum_attn.config``, ``quantum_attn.nn``, ``quantum_attn.ops``, ``quantum_attn.quantum_attn_interface``)."""
import sys

import quantumattention_b200 as _impl
from quantumattention_b200 import *  # noqa: F401,F403
from quantumattention_b200 import QuantizedKV, __all__, __version__, config, nn, ops, quantize_kv, quantum_attn_interface  # noqa: F401

for _name in ("config", "nn", "ops", "quantum_attn_interface"):
    sys.modules[f"{__name__}.{_name}"] = getattr(_impl, _name)
