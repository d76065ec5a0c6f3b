// acxrun — process launcher for acx-b200 programs (the mpiexec stand-in; the reference
// delegates launch to `mpiexec -np N`, /root/reference/README.md:99-103).
//
//   acxrun -np N [--timeout SEC] [--tag-output] [--bind-gpu] prog [args...]
//
// Forks N ranks with ACX_RANK / ACX_SIZE / ACX_LOCAL_RANK / ACX_JOBID set (the same
// variables torchrun's RANK/WORLD_SIZE/LOCAL_RANK map onto), forwards signals, kills
// the whole job when any rank fails or the timeout expires, and exits with the worst
// rank status. --bind-gpu sets ACX_DEVICE=<local rank> (ranks otherwise pick
// rank % ndevices themselves, as the reference tests do).
#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

static std::vector<pid_t> g_kids;
static void kill_all(int sig) {
    for (pid_t p : g_kids) if (p > 0) kill(p, sig);
}
static void on_signal(int sig) { kill_all(sig); }

static double now() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
    int np = 1, i = 1;
    double timeout = 0;
    bool tag = false, bind = false;
    for (; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-np" || a == "-n") && i + 1 < argc) np = atoi(argv[++i]);
        else if (a == "--timeout" && i + 1 < argc) timeout = atof(argv[++i]);
        else if (a == "--tag-output") tag = true;
        else if (a == "--bind-gpu") bind = true;
        else if (a == "-h" || a == "--help") {
            printf("usage: acxrun -np N [--timeout SEC] [--tag-output] [--bind-gpu] prog [args...]\n");
            return 0;
        } else break;
    }
    if (i >= argc || np < 1) { fprintf(stderr, "acxrun: usage: acxrun -np N prog [args...]\n"); return 2; }
    char jobid[64];
    snprintf(jobid, sizeof jobid, "%d_%lld", (int)getpid(), (long long)time(nullptr));
    signal(SIGINT, on_signal);
    signal(SIGTERM, on_signal);
    g_kids.assign(np, -1);
    for (int r = 0; r < np; r++) {
        int pfd[2] = {-1, -1};
        if (tag && pipe(pfd) != 0) { perror("pipe"); return 2; }
        pid_t pid = fork();
        if (pid < 0) { perror("fork"); kill_all(SIGKILL); return 2; }
        if (pid == 0) {
            char buf[32];
            snprintf(buf, sizeof buf, "%d", r);  setenv("ACX_RANK", buf, 1); setenv("ACX_LOCAL_RANK", buf, 1);
            if (bind) setenv("ACX_DEVICE", buf, 1);
            snprintf(buf, sizeof buf, "%d", np); setenv("ACX_SIZE", buf, 1);
            setenv("ACX_JOBID", jobid, 1);
            if (tag) {
                close(pfd[0]);
                // line-tagging is done by a tiny reader child per rank
                dup2(pfd[1], 1); dup2(pfd[1], 2); close(pfd[1]);
            }
            execvp(argv[i], &argv[i]);
            fprintf(stderr, "acxrun: exec %s failed: %s\n", argv[i], strerror(errno));
            _exit(127);
        }
        g_kids[r] = pid;
        if (tag) {
            close(pfd[1]);
            if (fork() == 0) {  // reader: prefix every line with [rank]
                FILE *in = fdopen(pfd[0], "r");
                char *line = nullptr; size_t cap = 0;
                while (getline(&line, &cap, in) > 0) { printf("[%d] %s", r, line); fflush(stdout); }
                _exit(0);
            }
            close(pfd[0]);
        }
    }
    int worst = 0, live = np;
    double t0 = now();
    bool killed = false;
    while (live > 0) {
        int st = 0;
        pid_t p = waitpid(-1, &st, WNOHANG);
        if (p > 0) {
            int r = -1;
            for (int k = 0; k < np; k++) if (g_kids[k] == p) r = k;
            if (r < 0) continue;  // a reader child
            g_kids[r] = -1; live--;
            int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
            if (code != 0) {
                if (!killed) fprintf(stderr, "acxrun: rank %d exited with status %d%s\n", r, code, live ? "; terminating job" : "");
                if (!killed && (worst == 0 || code > worst)) worst = code;  // ranks we terminate ourselves do not count
                if (!killed && live) { kill_all(SIGTERM); killed = true; t0 = now(); timeout = 5; }
            }
            continue;
        }
        if (timeout > 0 && now() - t0 > timeout) {
            if (!killed) { fprintf(stderr, "acxrun: timeout after %.0f s; killing job\n", timeout); worst = worst ? worst : 124; }
            kill_all(SIGKILL);
            killed = true; timeout = 0;
        }
        usleep(2000);
    }
    while (waitpid(-1, nullptr, 0) > 0) {}  // reap readers
    // best-effort removal of a segment left by a crashed rank 0
    std::string seg = std::string("/dev/shm/acx_") + jobid;
    unlink(seg.c_str());
    return worst;
}
