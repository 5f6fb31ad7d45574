# coding=utf-8
from . import drop_edge
