# coding=utf-8
from . import common_pool, set2set, topk_pool, score_pool
