// kernel-launch counter behind b200_launch_count() (bench.py reports it as gpu_launches)
#pragma once
#include <stdint.h>
void b200_count_launch();
