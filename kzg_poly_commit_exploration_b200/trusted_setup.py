"""Host mirror of the reference's `trusted_setup` module (src/trusted_setup.rs).

`SetupArtifactsGenerator(secret)` is the reference's iterator; here `.take(n)` runs ONE bulk call
(kzg_setup_generate: fixed-base kernel on the GPU) and returns a `TrustedSetup` that behaves like the
reference's `Vec<SetupArtifact>` for reading (len, indexing -> SetupArtifact with `.g1`) and carries
the device-resident bases the commit / open kernels use.  The G2 half of SetupArtifact
(src/trusted_setup.rs:64-72) is outside the accelerated path."""
from __future__ import annotations

from .curves import G1Point
from .engine import Context, Setup

_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class SetupArtifact:
    __slots__ = ("g1",)

    def __init__(self, g1: G1Point):
        self.g1 = g1

    def to_json(self):
        return {"g1": self.g1.to_json()}


class TrustedSetup:
    """Vec<SetupArtifact> + its device copy."""

    def __init__(self, ctx: Context, dev: Setup):
        self.ctx = ctx
        self.dev = dev
        self._g1_cache = None

    def __len__(self):
        return len(self.dev)

    def _g1(self):
        if self._g1_cache is None:
            self._g1_cache = self.dev.export_g1()
        return self._g1_cache

    def __getitem__(self, i):
        if isinstance(i, slice):
            pts = self._g1()
            idx = range(*i.indices(len(self)))
            sub = b"".join(pts[48 * k: 48 * k + 48] for k in idx)
            return TrustedSetup(self.ctx, self.ctx.setup_import_g1(sub))
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        pts = self._g1()
        return SetupArtifact(G1Point._raw(pts[48 * i: 48 * i + 48]))

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    @classmethod
    def from_artifacts(cls, artifacts, ctx: Context | None = None) -> "TrustedSetup":
        ctx = ctx or default_context()
        return cls(ctx, ctx.setup_import_g1(b"".join(a.g1.compress() for a in artifacts)))


class SetupArtifactsGenerator:
    """SetupArtifactsGenerator::new(secret) — `secret` is 32 big-endian bytes (src/trusted_setup.rs:20-28)."""

    def __init__(self, secret: bytes, ctx: Context | None = None):
        if len(secret) != 32:
            raise ValueError("secret must be 32 bytes")
        self.secret = bytes(secret)
        self.ctx = ctx or default_context()

    def take(self, n: int) -> TrustedSetup:
        """.take(n).collect(): items 0..n, item 0 = the generator (src/trusted_setup.rs:40-78)."""
        return TrustedSetup(self.ctx, self.ctx.setup_generate(self.secret, n))
