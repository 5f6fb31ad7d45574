# coding=utf-8
from .graph import Graph, BatchGraph
