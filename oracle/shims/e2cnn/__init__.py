"""Shim: inert e2cnn placeholder (only lets the reference's escnn_networks.py import)."""


class _Inert:
    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        raise RuntimeError("e2cnn is not installed in this image; ESCNN networks cannot be built")

    def __mro_entries__(self, bases):
        return (object,)


gspaces = _Inert()
nn = _Inert()
