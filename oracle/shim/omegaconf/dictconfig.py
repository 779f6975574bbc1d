class DictConfig(dict):
    pass
