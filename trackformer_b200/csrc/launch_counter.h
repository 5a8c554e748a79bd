// Process-wide count of kernel launches enqueued by this library (all .cu files); read by msda_b200_launch_count().
#pragma once
extern "C" void msda_b200_count_launches(int n);
