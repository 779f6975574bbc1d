from fatezero_b200.tables import get_refinement_mapper, get_replacement_mapper, get_word_inds  # noqa: F401
