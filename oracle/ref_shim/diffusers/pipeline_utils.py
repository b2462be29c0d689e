import torch
from tqdm.auto import tqdm


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self):
        for m in vars(self).values():
            if isinstance(m, torch.nn.Module):
                return next(m.parameters()).device
        return torch.device("cpu")

    def progress_bar(self, iterable=None, total=None):
        cfg = dict(getattr(self, "_progress_bar_config", {}))
        cfg.setdefault("disable", True)
        return tqdm(iterable, **cfg) if iterable is not None else tqdm(total=total, **cfg)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        return [Image.fromarray(image) for image in images]
