"""Configuration dataclasses with the reference's field names and defaults
(active_inference_diffusion/configs/config.py:10-86) for the fields the hot path reads."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class DiffusionConfig:
    num_diffusion_steps: int = 1000
    beta_start: float = 1e-4
    beta_end: float = 0.02
    beta_schedule: str = "cosine"  # "cosine" | "linear"
    prediction_type: str = "score"
    use_continuous_time: bool = True
    time_annealing_start: float = 1.0
    time_annealing_end: float = 0.1
    annealing_steps: int = 100000
    gradient_clip_val: float = 0.1


@dataclass
class BeliefDynamicsConfig:
    use_belief_dynamics: bool = True
    belief_dim: int = 50
    diffusion_coefficient: float = 0.1
    learning_rate: float = 0.1
    dt: float = 0.01
    min_variance: float = 1e-6
    max_variance: float = 10.0
    use_full_covariance: bool = False
    noise_scale: float = 0.01


@dataclass
class ActiveInferenceConfig:
    env_name: str = "HalfCheetah-v4"
    observation_dim: int = 17
    action_dim: int = 6
    precision_init: float = 1.0
    expected_free_energy_horizon: int = 5
    efe_horizon: int = 5
    epistemic_weight: float = 0.1
    extrinsic_weight: float = 1.0
    pragmatic_weight: float = 1.0
    consistency_weight: float = 0.1
    discount_factor: float = 0.99
    contrastive_weight: float = 0.5
    kl_weight: float = 0.1
    diffusion_weight: float = 1.0
    reward_weight: float = 0.5
    hidden_dim: int = 512
    latent_dim: int = 128
    spatial_aggregator_output_dim: int = 256
    num_layers: int = 3
    pixel_observation: bool = False
    batch_size: int = 256
    learning_rate: float = 5e-5
    gradient_clip: float = 0.5
    preference_temperature: float = 1.0
    preference_learning_rate: float = 0.01
    min_preference_temperature: float = 0.1
    max_preference_temperature: float = 10.0
    temperature_decay: float = 0.995
    use_reward_preferences: bool = True
    baseline_reward: float = 0.0
    preference_momentum: float = 0.9
    diffusion: DiffusionConfig = field(default_factory=DiffusionConfig)
    belief_dynamics: BeliefDynamicsConfig = field(default_factory=BeliefDynamicsConfig)
    device: str = "cuda"
