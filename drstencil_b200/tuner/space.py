"""Search space, validity filter and configuration names.

Reference: benchmarks/2d5pt_star/tuning.py:13-86,124-139 (2D) and benchmarks/3d7pt_star/tuning.py
(3D).  A configuration is the reference's tuple
  (step, dist, (bx, by), streaming, sn, s_unroll, blockMergeX, mx, blockMergeY, my, merge_forward, prefetch)
extended by the engine's own axes (ring stages, warps per CTA, launch-bounds blocks, vectors per
thread, rows per thread in 3D).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field, replace
from typing import Iterator, List

from .. import Knobs


@dataclass(frozen=True)
class Config:
    step: int = 1
    dist: int = 0
    bx: int = 64
    by: int = 1
    streaming: bool = True
    sn: int = 128
    s_unroll: int = 4
    block_merge_x: bool = True
    mx: int = 1
    block_merge_y: bool = False
    my: int = 1
    merge_forward: int = 5
    prefetch: bool = False
    # engine axes
    stages: int = 4
    min_blocks: int = 0
    rows_3d: int = 0
    dtype: str = "f64"
    fuse: str = "temporal"
    # warps of a CTA along x / y that share one input ring (single-step 3D sweep)
    share_x: int = 1
    share_y: int = 1
    dim: int = 2        # selects the name grammar (the reference has one per dimensionality)

    def knobs(self) -> Knobs:
        k = Knobs(step=self.step, sn=self.sn, bx=self.bx, by=self.by, streaming=int(self.streaming),
                  stream_unroll=self.s_unroll, merge_forward=self.merge_forward, dtype=self.dtype, fuse=self.fuse)
        if self.dist:
            k.set("dist", self.dist)
        if self.block_merge_x:
            k.set("block_merge_x", self.mx)
        else:
            k.set("cyclic_merge_x", self.mx)
        if self.block_merge_y:
            k.set("block_merge_y", self.my)
        else:
            k.set("cyclic_merge_y", self.my)
        if self.prefetch:
            k.set("prefetch", 1)
        k.set("stages", self.stages)
        if self.min_blocks:
            k.set("min_blocks", self.min_blocks)
        if self.rows_3d:
            k.set("rows_3d", self.rows_3d)
        if self.share_x * self.share_y > 1:
            k.set("share_x", self.share_x)
            k.set("share_y", self.share_y)
        return k


def cfg_to_string(c: Config) -> str:
    """The reference's result-file name grammars (2D: benchmarks/2d5pt_star/tuning.py:72-86,
    fu{step}d{dist}bx{bx}(sn{sn}u{unroll} | y{by})(bmx|cmx){m}[(bmy|cmy){m}]mf{t}[p]; 3D:
    benchmarks/3d7pt_star/tuning.py:57-72, fu{step}d{dist}bx{bx}y{by}sn{sn}u{unroll}(bmx|cmx){m}(bmy|cmy){m}mf{t}[p]), followed by the
    engine-only axes when they differ from the defaults: st{stages} mb{min_blocks} ry{rows} and the
    dtype/fuse tags."""
    if c.dim == 3:      # benchmarks/3d7pt_star/tuning.py:57-72: block shape, sn and unroll are always part of the name
        s = "fu%dd%dbx%dy%dsn%du%d" % (c.step, c.dist, c.bx, c.by, c.sn, c.s_unroll)
        s += ("bmx" if c.block_merge_x else "cmx") + str(c.mx)
        s += ("bmy" if c.block_merge_y else "cmy") + str(c.my)
        s += "mf%d" % c.merge_forward
        if c.prefetch:
            s += "p"
    else:
        if c.streaming:
            s = "fu%dd%dbx%dsn%du%d" % (c.step, c.dist, c.bx, c.sn, c.s_unroll)
        else:
            s = "fu%dd%dbx%dy%d" % (c.step, c.dist, c.bx, c.by)
        s += ("bmx" if c.block_merge_x else "cmx") + str(c.mx)
        if not c.streaming:
            s += ("bmy" if c.block_merge_y else "cmy") + str(c.my)
        s += "mf%d" % c.merge_forward
        if c.prefetch and c.streaming:
            s += "p"
    if c.stages != 4:
        s += "st%d" % c.stages
    if c.min_blocks:
        s += "mb%d" % c.min_blocks
    if c.rows_3d:
        s += "ry%d" % c.rows_3d
    if c.share_x * c.share_y > 1:
        s += "sx%dsy%d" % (c.share_x, c.share_y)
    if c.dtype != "f64":
        s += c.dtype
    if c.fuse != "temporal":
        s += "alg"
    return s


def cfg_from_string(name: str, dim: int) -> Config:
    """Inverse of cfg_to_string: the configuration a result-file name denotes.  presets.py states the BASELINE
    presets as such names, so that a preset IS a tuner record (profiles/r02_tune_*.json) and can be checked
    against it."""
    import re
    rest = name
    kw = dict(dim=dim)

    def take(pattern):
        nonlocal rest
        m = re.match(pattern, rest)
        if not m:
            return None
        rest = rest[m.end():]
        return m

    m = take(r"fu(\d+)d(\d+)bx(\d+)")
    if not m:
        raise ValueError("not a configuration name: %r" % name)
    kw.update(step=int(m.group(1)), dist=int(m.group(2)), bx=int(m.group(3)))
    if dim == 3:
        m = take(r"y(\d+)sn(\d+)u(\d+)")
        if not m:
            raise ValueError("not a 3D configuration name: %r" % name)
        kw.update(by=int(m.group(1)), sn=int(m.group(2)), s_unroll=int(m.group(3)), streaming=False)
    else:
        m = take(r"sn(\d+)u(\d+)")
        if m:
            kw.update(streaming=True, by=1, sn=int(m.group(1)), s_unroll=int(m.group(2)))
        else:
            m = take(r"y(\d+)")
            if not m:
                raise ValueError("not a 2D configuration name: %r" % name)
            kw.update(streaming=False, by=int(m.group(1)))
    m = take(r"(bmx|cmx)(\d+)")
    kw.update(block_merge_x=m.group(1) == "bmx", mx=int(m.group(2)))
    m = take(r"(bmy|cmy)(\d+)")
    if m:
        kw.update(block_merge_y=m.group(1) == "bmy", my=int(m.group(2)))
    m = take(r"mf(\d+)")
    kw.update(merge_forward=int(m.group(1)))
    if take(r"p(?![a-z])"):
        kw.update(prefetch=True)
    m = take(r"st(\d+)")
    if m:
        kw.update(stages=int(m.group(1)))
    m = take(r"mb(\d+)")
    if m:
        kw.update(min_blocks=int(m.group(1)))
    m = take(r"ry(\d+)")
    if m:
        kw.update(rows_3d=int(m.group(1)))
    m = take(r"sx(\d+)sy(\d+)")
    if m:
        kw.update(share_x=int(m.group(1)), share_y=int(m.group(2)))
    if take(r"f32"):
        kw.update(dtype="f32")
    if take(r"alg"):
        kw.update(fuse="algebraic")
    if rest:
        raise ValueError("trailing %r in configuration name %r" % (rest, name))
    c = Config(**kw)
    if cfg_to_string(c) != name:
        raise ValueError("configuration name %r does not round-trip (%r)" % (name, cfg_to_string(c)))
    return c


def cfg_to_command_line(c: Config) -> str:
    """Arguments for the `drstencil` CLI that reproduce this configuration (tuning.py:50-69)."""
    cmd = " --step %d --dist %d --bx %d" % (c.step, c.dist, c.bx)
    if c.dim == 3:
        cmd += " --by %d --sn %d --stream-unroll %d" % (c.by, c.sn, c.s_unroll)
    elif c.streaming:
        cmd += " --streaming --sn %d --stream-unroll %d" % (c.sn, c.s_unroll)
    else:
        cmd += " --by %d" % c.by
    cmd += (" --block-merge-y %d" if c.block_merge_y else " --cyclic-merge-y %d") % c.my
    cmd += (" --block-merge-x %d" if c.block_merge_x else " --cyclic-merge-x %d") % c.mx
    cmd += " --merge-forward %d" % c.merge_forward
    if c.prefetch and c.streaming:
        cmd += " --prefetch"
    if c.stages != 4:
        cmd += " --stages %d" % c.stages
    if c.min_blocks:
        cmd += " --min-blocks %d" % c.min_blocks
    if c.rows_3d:
        cmd += " --rows-3d %d" % c.rows_3d
    if c.share_x * c.share_y > 1:
        cmd += " --share-x %d --share-y %d" % (c.share_x, c.share_y)
    if c.dtype != "f64":
        cmd += " --dtype " + c.dtype
    if c.fuse != "temporal":
        cmd += " --fuse algebraic"
    return cmd


def filter_config(c: Config, dim: int, radius: int, esize: int = 8) -> bool:
    """Validity and pruning.  The reference rejects tiles that do not cover the halo and shared
    memory above 32 KiB (tuning.py:13-47); here the limits are the B200's (227 KiB per CTA, 64
    warps and 64K registers per SM) plus a resource model that drops configurations which cannot
    keep enough bytes in flight to cover HBM latency."""
    vec = 16 // esize
    warps = max(1, (c.bx if (dim == 2 and c.streaming) else c.bx * c.by) // 32)
    if warps > 16 or (c.bx * max(1, c.by)) % 32:
        return False
    if c.stages not in (2, 4, 8) or c.s_unroll not in (1, 2, 4, 8, 16):
        return False
    if dim == 2:
        vt = min(2, c.mx) if c.block_merge_x else 1
        cols = vt * vec
        e = radius
        hw = ((c.step - 1) * e + vec - 1) // vec * vec if c.fuse == "temporal" else 0
        if 32 * cols - 2 * hw < vec:                       # not covering the halo region
            return False
        wb = 32 * cols + 2 * ((e + vec - 1) // vec * vec)
        if wb > 256:
            return False
        stage = (c.s_unroll * wb * esize + 127) // 128 * 128
        smem = warps * c.stages * (stage + 8)
        if c.fuse == "temporal" and c.step > 1:      # scatter: partial sums per level + the row in flight
            live = (c.step * (2 * radius + 1) * cols + 2 * (cols + 2 * e)) * (esize // 4)
        else:
            live = (2 * radius + 1) * (cols + 2 * e) * (esize // 4)
    else:
        ry = c.rows_3d or 8
        wb = 32 * vec + 2 * ((radius + vec - 1) // vec * vec)
        stage = (wb * (ry + 2 * radius) * esize + 127) // 128 * 128
        if c.stages < 2 * radius + 2:
            return False
        smem = warps * c.stages * (stage + 8)
        if c.share_x * c.share_y > 1:                       # one ring per CTA (drs_sweep3d_cta.cuh)
            if c.share_x * c.share_y != warps or c.step != 1 and c.fuse == "temporal":
                return False
            wb = c.share_x * 32 * vec + 2 * ((radius + vec - 1) // vec * vec)
            if wb > 256:
                return False
            stage = (wb * (c.share_y * ry + 2 * radius) * esize + 127) // 128 * 128
            smem = c.stages * (stage + 16)
        live = (2 * radius + 1) * ry * vec * (esize // 4)
    if smem > 227 * 1024:
        return False
    if live + 40 > 255:                                     # the window cannot live in registers
        return False
    # bytes a resident SM keeps in flight: must cover ~44 KB (6.5 TB/s x ~1 us / 148 SMs)
    regs = min(255, live + 56)
    ctas = min(32, 227 * 1024 // max(smem, 1), 65536 // (regs * warps * 32))
    if c.min_blocks and c.min_blocks > ctas:
        return False
    rings = 1 if c.share_x * c.share_y > 1 else warps
    in_flight = ctas * rings * (c.stages - 1 if dim == 2 else c.stages - 2 * radius) * stage
    return ctas >= 1 and in_flight >= 24 * 1024


def search_space(dim: int, radius: int, step: int = 1, dtype: str = "f64", fuse: str = "temporal",
                 experimental: bool = True) -> List[Config]:
    """Cartesian product of the axes (reference: tuning.py:124-139), filtered.  For 3D single-step sweeps the
    space also holds six / eight / twelve rows per thread, longer chunks and the CTA-shared input ring
    (share_x x share_y; `experimental=False` leaves those out)."""
    esize = 8 if dtype == "f64" else 4
    out = []
    if dim == 2:
        for bx, sn, unroll, mx, stages, mb in itertools.product(
                (32, 64, 128), (32, 64, 128, 256, 512), (2, 4, 8), (1, 2), (2, 4), (0, 4)):
            c = Config(step=step, bx=bx, by=1, streaming=True, sn=sn, s_unroll=unroll, block_merge_x=True, mx=mx,
                       stages=stages, min_blocks=mb, dtype=dtype, fuse=fuse)
            if filter_config(c, 2, radius, esize):
                out.append(c)
    else:
        for (bx, by), sn, my, stages in itertools.product(
                ((32, 1), (32, 2), (32, 4)), (8, 16, 32, 64), (1, 2), (4, 8)):
            c = Config(step=step, bx=bx, by=by, streaming=False, sn=sn, block_merge_y=True, my=my, rows_3d=4 * my,
                       stages=stages, dtype=dtype, fuse=fuse, dim=3)
            if filter_config(c, 3, radius * (step if fuse == "algebraic" or dim == 3 else 1), esize):
                out.append(c)
        if experimental and step == 1:
            for (sx, sy), sn, ry, stages in itertools.product(
                    ((1, 1), (2, 1), (1, 2), (2, 2), (3, 2), (2, 4)), (32, 64, 128), (4, 6, 8, 12), (4, 8)):
                warps = sx * sy if sx * sy > 1 else 4
                c = Config(step=1, bx=32, by=warps, streaming=False, sn=sn, block_merge_y=True, my=1, rows_3d=ry,
                           stages=stages, dtype=dtype, fuse=fuse, share_x=sx, share_y=sy, dim=3)
                if filter_config(c, 3, radius, esize) and c not in out:
                    out.append(c)
    return out
