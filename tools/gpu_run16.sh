SCADE_TC_TRACE=1 timeout 300 python tools/tc_trace.py 2>&1 | tail -40
