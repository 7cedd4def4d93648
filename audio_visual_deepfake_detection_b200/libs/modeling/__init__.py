from .models import make_backbone, make_neck, make_meta_arch, make_generator
from . import meta_archs      # registers the two meta-archs, the backbone / neck descriptors and the point generator
from .spec import EXP5, EXP12, EXP13, state_dict_spec

__all__ = ["make_backbone", "make_neck", "make_meta_arch", "make_generator", "EXP5", "EXP12", "EXP13", "state_dict_spec"]
