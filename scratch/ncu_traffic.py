"""Write / update profiles/ncu_traffic.json from `ncu --set full` reports: per kernel, DRAM bytes of ONE launch
(dram__bytes_read.sum + dram__bytes_write.sum), which bench.py reports as roofline.traffic.
usage: python scratch/ncu_traffic.py <workload>:<photons> "<kernel label>"=report.ncu-rep ..."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def dram_bytes(report):
    txt = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, first = rows[0], rows[1], rows[2]
    total = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        total += float(first[i].replace(',', '')) * UNIT[units[i]]
    return int(total), first[hdr.index('Kernel Name')], float(first[hdr.index('gpu__time_duration.sum')].replace(',', ''))


def main():
    key = sys.argv[1]
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    table = json.load(open(path)) if os.path.exists(path) else {}
    entry = table.setdefault(key, {})
    for spec in sys.argv[2:]:
        label, report = spec.split('=', 1)
        b, kernel, dur = dram_bytes(report)
        entry[label] = {'dram_bytes': b, 'kernel_name': kernel, 'report': os.path.basename(report), 'duration_under_ncu': dur}
        print(label, b, kernel, dur)
    json.dump(table, open(path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
