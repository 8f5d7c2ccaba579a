from .constants import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD  # noqa: F401


def create_transform(*args, **kwargs):  # only imported by dataset code that is out of scope
    raise NotImplementedError("timm.data.create_transform is not part of the oracle shim")
