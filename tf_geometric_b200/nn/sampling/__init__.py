# coding=utf-8
from .drop_edge import drop_edge
