"""Exception type of the package (the reference's ZstdError, c-ext/backend_c.c:243-246)."""


class ZstdError(Exception):
    pass
