"""TEST INFRASTRUCTURE ONLY: every pyplot call is a no-op (the examples only plot at the end)."""


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __getitem__(self, item):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))


def __getattr__(name):
    return _Anything()
