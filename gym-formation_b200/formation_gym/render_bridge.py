"""Render bridge: draws ONE env's state on the host, like ``MultiAgentEnv.render`` of the reference
(formation_gym/environment.py:243-393; viewer geometry formation_gym/rendering.py:44-150,242-256): every entity is
a filled circle of radius ``entity.size`` in ``entity.color`` (agents half transparent, :297,:375), walls are filled
rectangles (:329-345), the camera is a square of half-width ``cam_range`` = 2 (environment.py:7) around the
agents' centroid (shared viewer, :363-369) on a 700 x 700 canvas (:280).

The reference draws through pyglet/OpenGL, which this image does not have; visualisation is not on the step path,
so the bridge is a small numpy rasteriser that needs neither a display nor the GPU: ``mode='rgb_array'`` returns the
frame(s), ``mode='human'`` shows the same frame in a pyglet window when pyglet is importable.
"""
import numpy as np

CAM_RANGE = 2.0          # formation_gym/environment.py:7
CANVAS = 700             # rendering.Viewer(700, 700), environment.py:280


def rasterize(circles, walls=(), center=(0.0, 0.0), cam_range=CAM_RANGE, size=CANVAS):
    """circles: iterable of (x, y, radius, (r, g, b), alpha); walls: iterable of ((x0, y0, x1, y1), (r, g, b), alpha).
    Returns a uint8 image [size, size, 3], row 0 at the TOP (like Viewer.render(return_rgb_array=True))."""
    img = np.ones((size, size, 3), np.float64)
    scale = size / (2.0 * cam_range)
    left, bottom = center[0] - cam_range, center[1] - cam_range

    def blend(ys, xs, mask, color, alpha):
        region = img[ys, xs]
        region[mask] = (1.0 - alpha) * region[mask] + alpha * np.asarray(color[:3], np.float64)

    for (x0, y0, x1, y1), color, alpha in walls:
        px0, px1 = sorted(((x0 - left) * scale, (x1 - left) * scale))
        py0, py1 = sorted(((y0 - bottom) * scale, (y1 - bottom) * scale))
        c0, c1 = int(max(0, np.floor(px0))), int(min(size, np.ceil(px1)))
        r0, r1 = int(max(0, np.floor(size - py1))), int(min(size, np.ceil(size - py0)))
        if c1 > c0 and r1 > r0:
            blend(slice(r0, r1), slice(c0, c1), np.ones((r1 - r0, c1 - c0), bool), color, alpha)
    for x, y, radius, color, alpha in circles:
        if not (np.isfinite(x) and np.isfinite(y)):
            continue                                     # NaN state (coincident agents, core.py:312): nothing to draw
        cx, cy, pr = (x - left) * scale, size - (y - bottom) * scale, max(radius * scale, 0.5)
        c0, c1 = int(max(0, np.floor(cx - pr))), int(min(size, np.ceil(cx + pr) + 1))
        r0, r1 = int(max(0, np.floor(cy - pr))), int(min(size, np.ceil(cy + pr) + 1))
        if c1 <= c0 or r1 <= r0:
            continue
        yy, xx = np.mgrid[r0:r1, c0:c1]
        mask = (xx + 0.5 - cx) ** 2 + (yy + 0.5 - cy) ** 2 <= pr * pr
        blend(slice(r0, r1), slice(c0, c1), mask, color, alpha)
    return (img * 255.0 + 0.5).astype(np.uint8)


def wall_rect(orient, axis_pos, end0, end1, width):
    """Corner rectangle of a core.Wall (environment.py:329-338)."""
    lo, hi = axis_pos - 0.5 * width, axis_pos + 0.5 * width
    return (end0, lo, end1, hi) if orient in ('H', 0) else (lo, end0, hi, end1)


def world_frame(world, agents=None, center=None, cam_range=CAM_RANGE, size=CANVAS):
    """One frame of a host-side ``World`` (the facade's records)."""
    circles = []
    for e in world.entities:
        color = (0.5, 0.5, 0.5) if e.color is None else tuple(np.asarray(e.color, np.float64)[:3])
        is_agent = 'agent' in e.name
        circles.append((float(e.state.p_pos[0]), float(e.state.p_pos[1]), float(e.size), color, 0.5 if is_agent else 1.0))
    walls = [(wall_rect(w.orient, float(w.axis_pos), float(w.endpoints[0]), float(w.endpoints[1]), float(w.width)),
              tuple(np.asarray(w.color, np.float64)[:3]), 1.0 if w.hard else 0.5) for w in world.walls]
    if center is None:
        P = np.stack([a.state.p_pos for a in (agents if agents is not None else world.agents)])
        center = np.nanmean(P, 0) if np.isfinite(P).any() else (0.0, 0.0)
    return rasterize(circles, walls, center, cam_range, size)


def show(frame, window=None):
    """``mode='human'``: blit the frame into a pyglet window (created on first use).  Raises ImportError when pyglet
    is missing -- use ``mode='rgb_array'`` on headless machines."""
    import pyglet
    h, w = frame.shape[:2]
    if window is None:
        window = pyglet.window.Window(width=w, height=h)
    window.switch_to()
    window.dispatch_events()
    window.clear()
    pyglet.image.ImageData(w, h, 'RGB', np.ascontiguousarray(frame[::-1]).tobytes()).blit(0, 0)
    window.flip()
    return window
