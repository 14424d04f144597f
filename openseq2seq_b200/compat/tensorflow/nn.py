from ._sym import relu  # noqa: F401
