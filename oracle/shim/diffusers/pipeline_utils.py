"""DiffusionPipeline restated (only what stable_diffusion.py / p2p_ddim_spatial_temporal.py touch)."""
import inspect
import torch
from .configuration_utils import ConfigMixin


class _Bar:
    def __init__(self, total=None): self.total = total
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def update(self, n=1): pass


class DiffusionPipeline(ConfigMixin):
    _optional_components = []

    def register_modules(self, **kwargs):
        for name, module in kwargs.items():
            self.register_to_config(**{name: (type(module).__module__, type(module).__name__)})
            setattr(self, name, module)

    @property
    def device(self):
        for name in self.config.keys():
            m = getattr(self, name, None)
            if isinstance(m, torch.nn.Module):
                return next(m.parameters()).device
        return torch.device("cpu")

    def to(self, device):
        for name in self.config.keys():
            m = getattr(self, name, None)
            if isinstance(m, torch.nn.Module):
                m.to(device)
        return self

    def progress_bar(self, iterable=None, total=None):
        return _Bar(total)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        return [Image.fromarray(image) for image in images]

    @staticmethod
    def _get_signature_keys(obj):
        parameters = inspect.signature(obj.__init__).parameters
        required = {k: v for k, v in parameters.items() if v.default is inspect._empty}
        optional = set({k for k, v in parameters.items() if v.default is not inspect._empty})
        return set(required.keys()) - set(["self"]), optional
