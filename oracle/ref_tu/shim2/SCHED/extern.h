/* TEST INFRASTRUCTURE: empty shadow of openair1/SCHED/extern.h */
