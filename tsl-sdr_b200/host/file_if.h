/* File IQ source (multifm/file_if.h in the reference): device.type == "file". */
#ifndef B200_FILE_IF_H
#define B200_FILE_IF_H
#include "receiver.h"

struct file_worker_thread;
aresult_t file_worker_thread_new(struct receiver **pthr, const jnode *cfg);
#endif
