from fatezero_b200.tables import get_time_words_attention_alpha, get_word_inds  # noqa: F401
