"""``NameDict``: insertion-ordered dict keyed by ``obj.name``.

Same surface as ``simudo/util/name_dict.py:7-84`` (iteration yields values, in
insertion order -- this fixes the band order of the mixed space,
``poisson_drift_diffusion.py:115-123``).
"""

__all__ = ['NameDict']

_NULL = object()


class NameDict(object):
    def __init__(self, iterable=None):
        self.mapping = {}
        if iterable is not None:
            for x in iterable:
                self.add(x)

    def __iter__(self):
        return iter(self.mapping.values())

    def __len__(self):
        return len(self.mapping)

    def __contains__(self, key):
        return key in self.mapping

    def __getitem__(self, key):
        return self.mapping[key]

    def __setitem__(self, key, value):
        self.replace(value)

    def obj_to_key(self, obj):
        return obj.name

    def add(self, obj, existing_raise=True):
        key = self.obj_to_key(obj)
        obj0 = self.mapping.get(key, _NULL)
        if obj0 is not _NULL:
            if existing_raise and obj0 is not obj:
                raise KeyError('key with different value {!r}'.format(key))
            return obj0
        self.mapping[key] = obj
        return obj

    def replace(self, obj):
        self.mapping[self.obj_to_key(obj)] = obj

    def items(self):
        return self.mapping.items()

    def keys(self):
        return self.mapping.keys()

    def values(self):
        return self.mapping.values()

    def update(self, data, existing_raise=True):
        if hasattr(data, 'values'):
            data = data.values()
        for x in data:
            self.add(x, existing_raise=existing_raise)
        return self

    def copy(self):
        return type(self)(self)

    def __or__(self, other):
        c = self.copy()
        c.update(other)
        return c

    def __repr__(self):
        return '{}({{{}}})'.format(type(self).__name__,
                                   ', '.join(repr(x) for x in self.mapping.values()))
