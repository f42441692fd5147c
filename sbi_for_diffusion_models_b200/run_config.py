"""Run configuration read by the hot path (mirrors the reference's run_config.py:4-43)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class RunConfig:
    # simulator
    MU_SENSORY: float = 1.0
    P_SUCCESS: float = 0.75
    # training set
    NUM_SIMULATIONS: int = 10_000
    TRAIN_BATCH_SIZE: int = 4096
    # observed session
    NUM_TRIALS_OBS: int = 50
    # x packing: log RT by hand, or let the density estimator do it
    LOG_RT_MANUALLY: bool = False
    SBI_LOG_TRANSFORM_X: bool = True
    Z_SCORE_X: Optional[str] = "independent"
    # MCMC
    NUM_CHAINS: int = 2
    WARMUP_STEPS: int = 100
    POSTERIOR_SAMPLES: int = 1000
    TEMPERATURE: float = 1.0
    THETA_TRUE_FROM_PRIOR: bool = True
    # SBC
    SBC_NUM_DATASETS: int = 10
    SBC_POST_SAMPLES: int = 1500


RUN_CONFIG_PARAMS = RunConfig()
