"""mirl_b200 -- B200-native (sm_100a) kernels behind MIRL's probabilistic-ensemble module API.

Public classes mirror the reference's (same constructor arguments, methods and return values):

====================  ===========================================  ============================
here                  reference class                              reference file
====================  ===========================================  ============================
B200PEModel           PEModel                                      mirl/agents/mbpo/pe_model.py
B200PEModelTrainer    PEModelTrainer                               mirl/agents/mbpo/pe_model_trainer.py
B200MBPOCollector     MBPOCollector                                mirl/agents/mbpo/mbpo_collector.py
B200MOPOCollector     MOPOCollector                                mirl/agents/mbpo/mopo_collector.py
B200EnsembleQValue    EnsembleQValue                               mirl/values/ensemble_value.py
B200TQCQValue         TQCQValue                                    mirl/values/distributional_value.py
B200DevicePool        SimplePool (imagined-data pool, in HBM)      mirl/pools/simple_pool.py
B200ExtraFieldPool    ExtraFieldPool (real-data pool, in HBM)      mirl/pools/extra_field_pool.py
B200CriticTrainer     critic half of SACAgent.train_from_torch_batch  mirl/agents/sac_agent.py:160-166
B200ActorTrainer      actor + alpha half of the same                  mirl/agents/sac_agent.py:170-195
B200SACUpdater        SACAgent.train_from_torch_batch                 mirl/agents/sac_agent.py:152-204
B200GaussianPolicy    GaussianPolicy (forward / sampling)           mirl/policies/gaussian_policy.py
====================  ===========================================  ============================

All computation goes through ``libmirl_b200.so`` (``include/mirl_b200.h``); there is no PyTorch
or CPU fallback -- calls raise if the library is missing or the tensors are not on a CUDA device.
"""
from .pe_model import B200PEModel                      # noqa: F401
from .pe_model_trainer import B200PEModelTrainer       # noqa: F401
from .values import B200CriticTrainer, B200EnsembleQValue, B200TQCQValue  # noqa: F401
from .collectors import B200MBPOCollector, B200MOPOCollector  # noqa: F401
from .pools import B200DevicePool, B200ExtraFieldPool  # noqa: F401
from .policies import B200ActorTrainer, B200GaussianPolicy  # noqa: F401
from .agents import B200SACUpdater                    # noqa: F401

__all__ = ["B200PEModel", "B200PEModelTrainer", "B200EnsembleQValue", "B200TQCQValue", "B200CriticTrainer",
           "B200MBPOCollector", "B200MOPOCollector", "B200DevicePool", "B200ExtraFieldPool", "B200GaussianPolicy",
           "B200ActorTrainer", "B200SACUpdater"]
