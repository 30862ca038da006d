"""Attribute-access dictionary (reference utils.py): `gan_synth_main.py` builds `spectral_params` and `hyper_params`
with it and the model code reads them as `hyper_params.generator_learning_rate`."""


class Dict(dict):
    """dict whose keys can also be read, written and deleted as attributes.  A missing key raises AttributeError on
    attribute reads (so `getattr(d, k, default)` and `hasattr` behave), KeyError on item reads as usual."""

    __slots__ = ()

    def __getattr__(self, key):
        if key in self:
            return dict.__getitem__(self, key)
        raise AttributeError("%s has no attribute or key %r" % (type(self).__name__, key))

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__
