"""Host mirror of the reference's `trusted_setup` module (src/trusted_setup.rs).

`SetupArtifactsGenerator(secret)` is the reference's iterator; here `.take(n)` runs ONE bulk call
(kzg_setup_generate: fixed-base kernel on the GPU) and returns a `TrustedSetup` that behaves like the
reference's `Vec<SetupArtifact>` for reading (len, indexing -> SetupArtifact with `.g1` and `.g2`) and
carries the device-resident bases the commit / open kernels use.  The G2 half of SetupArtifact
(src/trusted_setup.rs:64-72) comes from the same call sequence (kzg_setup_generate_g2, a fixed-base
kernel over Fq2): by default one per item like the reference; `take(n, g2_count=2)` keeps only the items
verification reads (setup[1].g2, src/polynomial.rs:284)."""
from __future__ import annotations

from .curves import G1Point, G2Point
from .engine import Context, Setup

_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class SetupArtifact:
    """SetupArtifact{g1, g2} (src/trusted_setup.rs:31-35); g2 is None for items generated without it."""
    __slots__ = ("g1", "g2")

    def __init__(self, g1: G1Point, g2: G2Point | None = None):
        self.g1 = g1
        self.g2 = g2

    def to_json(self):
        return {"g1": self.g1.to_json(), "g2": self.g2.to_json() if self.g2 is not None else None}

    @classmethod
    def from_json(cls, obj) -> "SetupArtifact":
        return cls(G1Point.from_json(obj["g1"]), G2Point.from_json(obj["g2"]) if obj.get("g2") is not None else None)


class TrustedSetup:
    """Vec<SetupArtifact> + its device copy."""

    def __init__(self, ctx: Context, dev: Setup):
        self.ctx = ctx
        self.dev = dev
        self._g1_cache = None
        self._g2_cache = None

    def _g2(self):
        if self._g2_cache is None:
            self._g2_cache = self.dev.export_g2()
        return self._g2_cache

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
            out = TrustedSetup(self.ctx, self.ctx.setup_import_g1(sub))
            g2 = self._g2()
            keep = [k for k in idx if 96 * k + 96 <= len(g2)]
            if keep and keep == list(idx)[: len(keep)]:
                out.dev.import_g2(b"".join(g2[96 * k: 96 * k + 96] for k in keep))
            return out
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        pts = self._g1()
        g2 = self._g2()
        return SetupArtifact(G1Point._raw(pts[48 * i: 48 * i + 48]),
                             G2Point._raw(g2[96 * i: 96 * i + 96]) if 96 * i + 96 <= len(g2) else None)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    @classmethod
    def from_artifacts(cls, artifacts, ctx: Context | None = None) -> "TrustedSetup":
        ctx = ctx or default_context()
        artifacts = list(artifacts)
        out = cls(ctx, ctx.setup_import_g1(b"".join(a.g1.compress() for a in artifacts)))
        g2 = []
        for a in artifacts:
            if a.g2 is None:
                break
            g2.append(a.g2.compress())
        if g2:
            out.dev.import_g2(b"".join(g2))
        return out


class SetupArtifactsGenerator:
    """SetupArtifactsGenerator::new(secret) — `secret` is 32 big-endian bytes (src/trusted_setup.rs:20-28)."""

    def __init__(self, secret: bytes, ctx: Context | None = None):
        if len(secret) != 32:
            raise ValueError("secret must be 32 bytes")
        self.secret = bytes(secret)
        self.ctx = ctx or default_context()

    def take(self, n: int, g2_count: int | None = None) -> TrustedSetup:
        """.take(n).collect(): items 0..n, item 0 = the generator pair (src/trusted_setup.rs:40-78)."""
        dev = self.ctx.setup_generate(self.secret, n)
        dev.generate_g2(self.secret, n if g2_count is None else min(g2_count, n))
        return TrustedSetup(self.ctx, dev)
