"""Room geometry -> 16-integer conditioning vector (the reference's rooms.py:5-99, consumed by the U-Net's
Embedding(2000, 256), u_net.py:257).

Vector layout: [a, b, c, d (wall lengths, cm), alpha, beta, gamma, delta (corner angles, deg), height,
                x_l, y_l, z_l (loudspeaker), x_m, y_m, z_m (microphone), rt60 (ms)], every entry rounded.
Loudspeaker l (1..60) sits on a 150 cm circle around the grid centre at angle (2l - 1) * pi / 60; microphone m sits in
one of five zones (A..E: centre shifted by -40 / +40 cm in x, +40 / -40 cm in y, or not at all) on a planar 8 x 8
grid with 4 cm pitch or on concentric circles of 30 microphones (radius 12 cm shrinking by 2 cm per ring).
Checked element for element against the reference's own rooms.py (tests/golden/rooms_golden.json)."""
from __future__ import annotations

import math

# dataset.py:84-89 -- the UTS rooms: four wall lengths, four corner angles, height, grid centre, RT60 in ms
UTS_ROOMS = {
    "AnechoicRoom": (490, 722, 490, 722, 90, 90, 90, 90, 529, (245, 361), 45),
    "HemiAnechoicRoom": (490, 722, 490, 722, 90, 90, 90, 90, 529, (245, 361), 52),
    "SmallMeetingRoom": (355, 410, 401, 378, 96, 90, 85, 88, 300, (175.5, 205), 497),
    "MediumMeetingRoom": (736, 520, 650, 434.5, 81, 92, 98, 89, 300, (368, 217.5), 659),
    "LargeMeetingRoom": (994, 923, 1087, 1022, 81.4, 105, 81.3, 92.3, 300, (497, 486.25), 1281),
    "ShoeBoxRoom": (600, 1175, 600, 1175, 90, 90, 90, 90, 300, (300, 881.25), 667),
}
_ZONE_SHIFT = {"A": (-40, 0), "B": (40, 0), "C": (0, 40), "D": (0, -40), "E": (0, 0)}
Z_PLANE = 145          # loudspeakers and microphones share one height (cm)


class UTSRoom:
    def __init__(self, a, b, c, d, alpha, beta, gamma, delta, height, grid_center, rt60):
        self.sides, self.angles, self.height = (a, b, c, d), (alpha, beta, gamma, delta), height
        self.grid_center, self.rt60 = tuple(grid_center), rt60

    def return_vector(self):
        return [round(v) for v in (*self.sides, *self.angles, self.height)]

    def get_m_l_position(self, characteristics):
        zone, array = characteristics[1], characteristics[2]
        l, m = int(characteristics[3]), int(characteristics[4])
        cx, cy = self.grid_center
        theta = (2 * l - 1) * math.pi / 60
        xl = round(-150 * math.sin(theta)) + cx
        yl = round(150 * math.cos(theta)) + cy
        xm = ym = 0
        if zone in _ZONE_SHIFT and array in ("Planar", "Circular"):
            sx, sy = _ZONE_SHIFT[zone]
            if array == "Planar":
                dx, dy = -14 + 4 * ((m - 1) % 8), 14 - 4 * math.floor((m - 1) / 8)
            else:
                radius = 12 - 2 * math.floor((m - 1) / 30)
                phi = ((m - 1) % 30) * 2 * math.pi / 30
                dx, dy = -radius * math.sin(phi), radius * math.cos(phi)
            xm, ym = dx + sx + cx, dy + sy + cy
        return [round(xl), round(yl), round(Z_PLANE), round(xm), round(ym), round(Z_PLANE), self.rt60]

    def return_embedding(self, characteristics):
        return self.return_vector() + self.get_m_l_position(characteristics)


def uts_room(name):
    return UTSRoom(*UTS_ROOMS[name])


def return_room(emb):
    """short room name from the first wall length of an embedding (rooms.py:101-115)."""
    return {490: "Anechoic", 355: "Small", 736: "Medium", 994: "Large", 600: "Box"}.get(emb[0])
