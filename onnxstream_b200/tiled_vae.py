"""Batched tiled VAE decode -- host-side mirror of sd.cpp's tiled decoders
(src/sd.cpp:1258-1346 `sd_tiled_decoder`, src/sd.cpp:2399-2503 the SDXL / any-size variant).

The reference decodes a large latent as overlapping 32x32 latent tiles (stride 24, the last tile of a row / column clamped to the edge),
each through one Model::run of a 32x32 -> 256x256 VAE decoder, and feather-blends tile k over the canvas: inside the 64-pixel
band at a tile's top / left edge (when it has a neighbour there) the new tile's weight ramps y/64 (x/64) from 0 to 1.  Tiles are
independent, so here ALL of them are pushed as batch siblings of ONE run (push_tensor semantics: same name pushed again = next
sibling, src/onnxstream.cpp:3040-3050): every decoder weight is fetched (streamed through the HBM ring) once for the whole image
instead of once per tile, and the engine walks the siblings inside each node.  The blend itself is the reference's float formula,
applied in the reference's tile order.

Works with any library exporting the reference's C ABI (the engine or the oracle): `model` is an onnxstream_b200.model.Model whose
graph maps `input_name` (1,4,T,T) -> `output_name` (1,3,8T,8T)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

TILE = 32       # latent tile edge (src/sd.cpp:1265, 2404)
STRIDE = 24     # tile pitch (src/sd.cpp:1325, 2480)
RAMP = 64       # feather width in pixels (src/sd.cpp:1310-1313)


def tile_origins(lat: int, tile: int = TILE, stride: int = STRIDE) -> List[int]:
    """Origins along one axis exactly as the reference's loops produce them: 0, 24, 48, ... with the last one clamped to lat - tile."""
    if lat < tile:
        raise ValueError("tiled decoder: resolution too small (use the untiled decoder)")
    out, v = [], 0
    while True:
        if v + tile > lat:
            v = lat - tile
        out.append(v)
        if v == lat - tile:
            return out
        v += stride


def blend_tile(canvas: np.ndarray, tile_img: np.ndarray, dx: int, dy: int, ramp: int = RAMP) -> None:
    """canvas[3,H,W] <- feather blend of tile_img[3,h,w] at (dx, dy): d = s*f + d*(1-f), f = (y/ramp if dy and y < ramp) * (x/ramp if dx and x < ramp)."""
    _, h, w = tile_img.shape
    fy = np.ones(h, np.float32)
    fx = np.ones(w, np.float32)
    if dy:
        fy[:ramp] = np.arange(ramp, dtype=np.float32) / np.float32(ramp)
    if dx:
        fx[:ramp] = np.arange(ramp, dtype=np.float32) / np.float32(ramp)
    f = (fy[:, None] * fx[None, :]).astype(np.float32)
    region = canvas[:, dy:dy + h, dx:dx + w]
    region[...] = tile_img * f + region * (np.float32(1) - f)


def tiled_decode(model, latent: np.ndarray, input_name: str, output_name: str, batched: bool = True, tile: int = TILE, stride: int = STRIDE,
                 upscale: int = 0) -> Tuple[np.ndarray, int]:
    """latent [1,4,H,W] float32 -> image [1,3,8H,8W]; returns (image, number of tiles).  batched=False decodes tile by tile (one run
    each: the reference's order of execution when its coroutine scheduler holds a single sample)."""
    latent = np.asarray(latent, np.float32)
    _, c, lh, lw = latent.shape
    ys, xs = tile_origins(lh, tile, stride), tile_origins(lw, tile, stride)
    origins = [(x, y) for y in ys for x in xs]
    tiles = [np.ascontiguousarray(latent[:, :, y:y + tile, x:x + tile]) for (x, y) in origins]
    outs = []
    if batched:
        model.clear_tensors()
        for t in tiles:
            model.add_tensor(input_name, t)
        model.run()
        for k in range(len(tiles)):
            o = model.get_tensor(output_name, k)
            if o is None:
                raise RuntimeError(f"tiled decode: output sibling {k} missing")
            outs.append(o)
    else:
        for t in tiles:
            model.clear_tensors()
            model.add_tensor(input_name, t)
            model.run()
            outs.append(model.get_tensor(output_name))
    upscale = int(outs[0].shape[-1]) // tile      # 8 for the SD decoders (three 2x upsamplers); read off the graph's own output
    canvas = np.zeros((3, lh * upscale, lw * upscale), np.float32)
    for (x, y), o in zip(origins, outs):
        blend_tile(canvas, o.reshape(3, tile * upscale, tile * upscale), x * upscale, y * upscale, (tile * upscale) // 4)     # 64 of 256 pixels in the reference
    return canvas[None], len(tiles)
