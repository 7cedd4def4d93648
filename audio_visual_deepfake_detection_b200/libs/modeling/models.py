"""Name -> class registries and builders, same surface as the reference's
libs/modeling/models.py:4-50 (register_backbone / register_neck /
register_generator / register_meta_arch and the four make_* builders).

On the accelerated path the backbone, neck and point generator are not
separate nn.Modules (they are stages of one kernel sequence, see engine.py);
their registered entries are descriptors that validate the configuration the
reference would have built, so `make_backbone(name, **kw)` etc. keep working
as configuration checks and unknown names fail exactly like the reference
(KeyError).
"""

backbones = {}
necks = {}
generators = {}
meta_archs = {}


def register_backbone(name):
    def decorator(cls):
        backbones[name] = cls
        return cls
    return decorator


def register_neck(name):
    def decorator(cls):
        necks[name] = cls
        return cls
    return decorator


def register_generator(name):
    def decorator(cls):
        generators[name] = cls
        return cls
    return decorator


def register_meta_arch(name):
    def decorator(cls):
        meta_archs[name] = cls
        return cls
    return decorator


def make_backbone(name, **kwargs):
    return backbones[name](**kwargs)


def make_neck(name, **kwargs):
    return necks[name](**kwargs)


def make_meta_arch(name, **kwargs):
    return meta_archs[name](**kwargs)


def make_generator(name, **kwargs):
    return generators[name](**kwargs)
