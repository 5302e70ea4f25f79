"""Model registry with the reference's lookup contract: ``MONO.module_dict[cfg.model['name']](cfg.model)``
(reference ``mono/model/registry.py:24-42``, used at ``train.py:79-81``)."""
import torch.nn as nn


class Registry:
    def __init__(self, name):
        self._name, self._module_dict = name, {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def register_module(self, cls):
        if not (isinstance(cls, type) and issubclass(cls, nn.Module)):
            raise TypeError("module must be a child of nn.Module, but got {}".format(cls))
        if cls.__name__ in self._module_dict:
            raise KeyError("{} is already registered in {}".format(cls.__name__, self._name))
        self._module_dict[cls.__name__] = cls
        return cls


MONO = Registry("mono")
