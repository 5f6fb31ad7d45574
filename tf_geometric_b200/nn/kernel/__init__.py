# coding=utf-8
from . import map_reduce, segment
