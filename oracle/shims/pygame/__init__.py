"""Headless no-op stub of the slice of pygame 1.9.4 the reference touches (game.py:40-49,155-229).
TEST INFRASTRUCTURE ONLY: it lets the unmodified reference ShipGame run without SDL.  Rendering, the
event queue and `clock.tick` sleeping (game.py:195) have no effect on the transition being pinned."""
import types

QUIT = 12
KEYDOWN = 2
K_ESCAPE, K_q, K_w, K_s, K_a, K_d = 27, 113, 119, 115, 97, 100


class _Surface(object):
    def __init__(self, size=(0, 0)):
        self.size = tuple(size)

    def fill(self, color):
        pass

    def get_size(self):
        return self.size


class _Clock(object):
    def tick(self, fps=0):
        return 0


def init():
    return (6, 0)


display = types.SimpleNamespace(set_mode=lambda size, *a, **k: _Surface(size),
                                set_caption=lambda *a, **k: None,
                                flip=lambda: None)
time = types.SimpleNamespace(Clock=_Clock)
key = types.SimpleNamespace(set_repeat=lambda *a, **k: None)
event = types.SimpleNamespace(get=lambda: [])
draw = types.SimpleNamespace(circle=lambda *a, **k: None, line=lambda *a, **k: None,
                             polygon=lambda *a, **k: None)
color = types.SimpleNamespace(THECOLORS={"white": (255, 255, 255, 255), "black": (0, 0, 0, 255),
                                         "green": (0, 255, 0, 255), "red": (255, 0, 0, 255)})
surfarray = types.SimpleNamespace(array3d=lambda surf: None)
