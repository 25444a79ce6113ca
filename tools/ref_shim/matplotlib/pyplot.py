"""matplotlib.pyplot stand-in for running the unmodified reference here: plotting is not part of any comparison, every call is a no-op."""


class _Anything:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Anything()
