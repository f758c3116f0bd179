from .inference import WanVideoGenerator

__all__ = ["WanVideoGenerator"]
