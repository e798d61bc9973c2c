"""CPU check of the noise-draw-counter protocol of the fused step (csrc/api.cu enqueue_step, csrc/integrator.cu k_integrate).

Every Langevin half step (O stage) must use draw index 0, 1, 2, ... in order, whatever mix of entry points produced it:
  * general launch sequence: the last block of a launch with an O stage writes counter + 1 (a "ticket");
  * ticketless iteration (one handle owning every bead, or a bead shard on the early-halo path): the opening O reads
    counter + 0 and leaves the counter alone, the kick-and-drift launch between the two thermostat launches -- it has no O
    stage, hence no reader of the counter -- adds 2 in its prologue, the closing O reads counter - 1.
A captured iteration is replayed with the offsets frozen, so the protocol has to hold for any counter value at entry."""
import itertools


class Device:
    def __init__(self):
        self.counter = 0
        self.draws = []

    def launch(self, o_stage, draw_off=0, bump=0, ticket=True):
        if o_stage:
            assert bump == 0
            self.draws.append(self.counter + draw_off)
            if ticket:
                assert draw_off == 0
                self.counter += 1
        else:
            self.counter += bump


def general_step(d):
    d.launch(True)                     # [SUBCM | O | SUM]
    d.launch(False)                    # [SUBCM | B | A]
    d.launch(True)                     # [assemble | B | O | SUM]


def ticketless_step(d):
    d.launch(True, draw_off=0, ticket=False)
    d.launch(False, bump=2)
    d.launch(True, draw_off=-1, ticket=False)


def piecewise_thermostat(d):
    d.launch(True)                     # pimdb_thermostat_step


def test_draw_indices_are_consecutive_for_every_mix_of_entry_points():
    ops = [general_step, ticketless_step, piecewise_thermostat]
    for seq in itertools.product(ops, repeat=5):
        d = Device()
        for op in seq:
            op(d)
        assert d.draws == list(range(len(d.draws))), [f.__name__ for f in seq]
        assert d.counter == len(d.draws)
