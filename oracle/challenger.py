"""Fiat-Shamir challenger (overwrite-mode duplex sponge over Poseidon-GL, rate 8).

TEST INFRASTRUCTURE.  Restates gnark-plonky2-verifier/challenger/challenger.go:42-49 (observe clears
output buffer, duplex when 8 inputs buffered), :89-98 (get_challenge pops from the END of the output
buffer), :146-165 (duplexing overwrites state[0..len), permutes, output = state[0..8)).
"""
from . import poseidon_gl as pg
from .gl import P

class Challenger:
    def __init__(self):
        self.state = [0] * 12
        self.inp = []
        self.out = []
    def observe(self, x):
        self.out = []
        self.inp.append(x % P)
        if len(self.inp) == pg.RATE:
            self._duplex()
    def observe_many(self, xs):
        for x in xs: self.observe(x)
    def observe_ext(self, e):
        self.observe(e[0]); self.observe(e[1])
    def observe_ext_many(self, es):
        for e in es: self.observe_ext(e)
    def _duplex(self):
        for i, x in enumerate(self.inp):
            self.state[i] = x
        self.inp = []
        self.state = pg.permute(self.state)
        self.out = list(self.state[:pg.RATE])
    def get_challenge(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()
    def get_n(self, n):
        return [self.get_challenge() for _ in range(n)]
    def get_ext(self):
        a = self.get_challenge(); b = self.get_challenge()
        return (a, b)
