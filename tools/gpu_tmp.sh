timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf8 or golden or well_formed or bitplane or config1 or config2" 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print('value',b['value'],'e2e',b['e2e']['value'])
for k in ('config1_validate_utf8_ascii_1GiB','validate_utf8_mixed_1GiB','next_detect_encodings_utf16_text','next_to_well_formed_utf16le'):
    print(k,b['extra'][k])
"
