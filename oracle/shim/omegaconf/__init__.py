"""tests-only stub: ptp_utils.get_time_words_attention_alpha only needs omegaconf.dictconfig.DictConfig to exist."""
from . import dictconfig
