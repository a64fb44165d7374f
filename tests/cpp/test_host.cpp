// tests/cpp/test_host.cpp -- replays the reference's integration test (/root/reference/test/index.coffee:12-117)
// through the C++ mirror of its Manager / Consumer / OpticalFlow classes (tidalwave_host.hpp) on top of the C ABI.
// usage: test_host <numThreads> <expect.pgm> <target.pgm> [<expect.pgm> <target.pgm> ...]
// Prints one JSON line per response / error, then the report.
#include "tidalwave_host.hpp"

using namespace tidalwave;

struct Printer : public Observer<Response, std::string, Report> {
    std::mutex mu;
    void onNext(const Response &r)
    {
        std::lock_guard<std::mutex> lk(mu);
        printf("{\"event\":\"data\",\"status\":\"%s\",\"span\":%d,\"threshold\":%g,\"expect_image\":\"%s\",\"target_image\":\"%s\",\"height\":%d,\"width\":%d,\"vector\":[",
               r.status.c_str(), r.span, r.threshold, r.expect_image.c_str(), r.target_image.c_str(), r.height, r.width);
        for (size_t i = 0; i < r.vectors.size(); i++)
            printf("%s{\"x\":%d,\"y\":%d,\"dx\":%.17g,\"dy\":%.17g}", i ? "," : "", r.vectors[i].x, r.vectors[i].y, r.vectors[i].dx, r.vectors[i].dy);
        printf("]}\n");
    }
    void onError(const std::string &reason)
    {
        std::lock_guard<std::mutex> lk(mu);
        printf("{\"event\":\"error\",\"status\":\"ERROR\",\"reason\":\"%s\"}\n", reason.c_str());
    }
    void onCompleted(const Report &rep)
    {
        std::lock_guard<std::mutex> lk(mu);
        printf("{\"event\":\"finish\",\"request\":%d,\"data\":%d,\"error\":%d}\n", rep.requestCount, rep.dataCount, rep.errorCount);
    }
};

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    Printer printer;
    Manager manager(&printer);
    Parameter p;
    p.threshold = 5.0; p.span = 10; p.numThreads = atoi(argv[1]); // defaults of src/broker.cpp:106-117
    p.optParam.pyrScale = 0.5; p.optParam.pyrLevels = 3; p.optParam.winSize = 30; p.optParam.pyrIterations = 3;
    p.optParam.polyN = 7; p.optParam.polySigma = 1.5; p.optParam.flags = 256;
    manager.start(p);
    for (int i = 2; i + 1 < argc; i += 2) manager.request(argv[i], argv[i + 1]);
    manager.waitAllAnswered();
    manager.stop();
    manager.join();
    return 0;
}
