"""
effective_model_learning_b200 -- B200-native evaluation of Lindblad TLS models for the
reversible-jump MCMC of henry104413/effective_model_learning.

Host-side classes keep the reference's names (BasicModel, LearningModel, LearningChain,
ProcessHandler, ParamsHandler, TwoLevelSystem, definitions.ops, configs); the dynamics and
the loss run in hand-written sm_100a CUDA kernels behind the C ABI of include/eml_b200.h.
"""
from . import (definitions, two_level_system, engine, basic_model, learning_model, params_handling,  # noqa: F401
               process_handling, learning_chain, configs, synthetic, dist, chains, predictive, output)
from .basic_model import BasicModel  # noqa: F401
from .learning_model import LearningModel  # noqa: F401
from .learning_chain import LearningChain  # noqa: F401
from .params_handling import ParamsHandler  # noqa: F401
from .process_handling import ProcessHandler  # noqa: F401
from .two_level_system import TwoLevelSystem  # noqa: F401
from .chains import BatchedChains  # noqa: F401

__version__ = '0.1.0'
