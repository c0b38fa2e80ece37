"""Palette and stroke constants of the MAGICAL world.

Restates the values of the reference `magical/style.py:10-40` (HLS lighten /
darken, the five Berkeley-derived base colours, line thicknesses and the arena
zoom-out factor).  Everything here is plain host-side math evaluated once at
import; the CUDA rasteriser only ever sees the resulting u8 triples.
"""
import colorsys

GOAL_LINE_THICKNESS = 0.01
SHAPE_LINE_THICKNESS = 0.015
ROBOT_LINE_THICKNESS = 0.01
# allocentric view shows [-ZOOM, ZOOM]^2 of the [-1, 1]^2 arena (style.py:40)
ARENA_ZOOM_OUT = 1.02


def rgb(r, g, b):
    """8-bit triple -> unit floats."""
    return (r / 255.0, g / 255.0, b / 255.0)


def _with_lightness(colour, fn):
    hue, light, sat = colorsys.rgb_to_hls(*colour)
    return colorsys.hls_to_rgb(hue, fn(light), sat)


def darken_rgb(colour):
    """HLS lightness x0.9 (reference style.py:10-14)."""
    return _with_lightness(colour, lambda l: max(0, l * 0.9))


def lighten_rgb(colour, times=1):
    """Move HLS lightness towards 1 by a factor 1.4**times (style.py:17-22)."""
    shrink = 1.4 ** times
    return _with_lightness(colour, lambda l: 1 - (1 - l) / shrink)


_BASE_HEX = {
    'blue': (0x3B, 0x7E, 0xA1),
    'yellow': (0xFD, 0xB5, 0x15),
    'red': (0xEE, 0x1F, 0x60),
    'green': (0x85, 0x94, 0x38),
}
COLOURS_RGB = {name: lighten_rgb(rgb(*hexv), 1.7)
               for name, hexv in _BASE_HEX.items()}
COLOURS_RGB['grey'] = rgb(162, 163, 175)
COLOURS_RGB['brown'] = rgb(224, 171, 118)

# clear colour of both views (reference base_env.py:186)
BACKGROUND_RGB = lighten_rgb(COLOURS_RGB['grey'], times=4)


def to_u8(colour):
    """Float colour -> the u8 triple a GL_RGB/UNSIGNED_BYTE framebuffer stores
    (round-to-nearest of 255*c, which is what GL's unorm conversion does)."""
    return tuple(int(round(255.0 * min(1.0, max(0.0, c)))) for c in colour)
