"""Dispatch of multivector methods on the (static) subspaces of their arguments.

Same contract as the reference's `numga/dynamic_dispatch.py:3-54`: implementations are tried in
registration order, the first whose predicate accepts the argument subspaces wins and is cached
per subspace tuple; `register(condition, position=0)` lets a user put an override in front
(the mechanism `numga/multivector/extension/optimized.py:24-33` uses).
"""
from types import MethodType


class SubspaceDispatch:
    def __init__(self, doc=''):
        self.__doc__ = doc
        self.options = []
        self.cache = {}

    def resolve(self, subspaces):
        try:
            return self.cache[subspaces]
        except KeyError:
            pass
        for condition, func in self.options:
            if condition(*subspaces):
                self.cache[subspaces] = func
                return func
        raise Exception('No matching implementation found')

    def __call__(self, *args):
        return self.resolve(tuple(a.subspace for a in args))(*args)

    def __get__(self, instance, owner):
        return self if instance is None else MethodType(self, instance)

    def register(self, condition=lambda *s: True, position=None):
        def wrap(func):
            self.options.insert(len(self.options) if position is None else position, (condition, func))
            self.cache.clear()
            return func
        return wrap
