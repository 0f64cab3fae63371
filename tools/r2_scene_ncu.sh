ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"Scene" -c 60 --csv --log-file /tmp/scene_l.csv python tools/bench_scene.py > /dev/null 2>&1
python tools/summarize_launches.py /tmp/scene_l.csv
python tools/bench_scene.py | cut -c1-330
