"""B200-native (sm_100a) implementation of the data-parallel hot path of
neuronphysics/active-inference-diffusion: reverse-diffusion sampling of the latent score
network, the score-matching/ELBO loss that trains it and EFE scoring of candidates, behind the
reference's own Python call surface.  See DESIGN.md."""
from .configs import ActiveInferenceConfig, BeliefDynamicsConfig, DiffusionConfig
from .diffusion import LatentDiffusionProcess
from .score_network import LatentScoreNetwork
from .heads import DiffusionConditionedPolicy, LatentDynamicsModel, ValueNetwork
from .active_inference import DiffusionActiveInference
from .pipeline import CandidateScorer
from .belief_dynamics import BeliefDynamics, FreeEnergyComputation
from .visual_encoder import DrQV2Encoder, SpatialAttention
from .train_utils import EMAModel, RunningMeanStd, update_belief_batched
from .train_graph import GraphedElboStep

__all__ = ["ActiveInferenceConfig", "BeliefDynamicsConfig", "DiffusionConfig",
           "LatentDiffusionProcess", "LatentScoreNetwork", "DiffusionConditionedPolicy",
           "LatentDynamicsModel", "ValueNetwork", "DiffusionActiveInference", "CandidateScorer",
           "BeliefDynamics", "FreeEnergyComputation", "DrQV2Encoder", "SpatialAttention", "RunningMeanStd",
           "update_belief_batched", "GraphedElboStep", "EMAModel"]
