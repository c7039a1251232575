"""trace_query.py for a chosen kernel variant: python profiles/trace_impl.py <impl> [bf16|bf16x3]"""
import runpy
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn

dfn.lib.dfn_debug_set_impl(int(sys.argv[1]))
sys.argv = ['trace_query.py'] + sys.argv[2:]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'trace_query.py'), run_name='__main__')
