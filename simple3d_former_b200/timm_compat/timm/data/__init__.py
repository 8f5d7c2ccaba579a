from .constants import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD  # noqa: F401


def create_transform(*args, **kwargs):
    raise NotImplementedError("image transforms are outside the hot path")
