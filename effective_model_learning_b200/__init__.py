"""
effective_model_learning_b200 -- B200-native evaluation of Lindblad TLS models for the
reversible-jump MCMC of henry104413/effective_model_learning.

Host-side classes keep the reference's names (BasicModel, LearningModel, LearningChain,
ProcessHandler, ParamsHandler, TwoLevelSystem, definitions.ops, configs); the dynamics and
the loss run in hand-written sm_100a CUDA kernels behind the C ABI of include/eml_b200.h.
"""
from . import definitions, two_level_system, engine, basic_model  # noqa: F401
from .basic_model import BasicModel  # noqa: F401
from .two_level_system import TwoLevelSystem  # noqa: F401

__version__ = '0.1.0'
