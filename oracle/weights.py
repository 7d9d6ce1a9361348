"""TEST INFRASTRUCTURE ONLY -- re-export of the synthetic weight / audio generator
(``libreasr_b200/synth.py``) so that oracle code and tests share one definition."""
from libreasr_b200.synth import *  # noqa: F401,F403
from libreasr_b200.synth import (BENCH_AUDIO_SEED, CONFIGS, LM_CONFIGS, LMConfig, ModelConfig, make_audio,  # noqa: F401
                                 make_lm_state_dict, make_state_dict)
