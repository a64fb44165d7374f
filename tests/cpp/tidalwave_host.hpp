// tidalwave_host.hpp -- TEST SCAFFOLDING (not product code, not a component): a REFERENCE-DERIVED host above the C ABI.
// It re-creates the reference's C++ operator / worker / manager class shapes (same names, argument meaning and error
// behaviour, minus OpenCV / libuv / V8, which are absent in this image) so that tests/cpp/test_host.cpp can replay
// /root/reference/test/index.coffee against libtidalwave_b200.so the way a maintainer's integration would call it.
// The product's dispatcher is tidal-wave_b200/csrc/tw_pool.cpp; nothing in the library includes this file.
//
//   OpticalFlowParameter, OpticalFlowStatus, ErrorCode, OpticalFlow::calculate / calculateInternal
//                                              /root/reference/src/opticalflow.h:9-52, src/opticalflow.cpp:20-94
//   Request, Vector, Response, Report, MessageQueue<T>         /root/reference/src/message_queue.h:13-118
//   Consumer (worker thread)                                   /root/reference/src/consumer.cpp:12-105
//   Parameter, Manager (start / request / stop / work)          /root/reference/src/manager.h:11-37, src/manager.cpp:40-98
//   Observer<Response, std::string, Report>                     /root/reference/src/observer.h:10-18
//
// cv::Mat is replaced by two plain containers (Image = CV_8UC1, Plane = CV_32FC1); cv::imread by imread_gray()
// (PNG, JPEG and binary PGM via tw_decode_gray -- SURVEY row f-1).  uv_thread / uv_mutex / uv_cond become std::thread /
// std::mutex / std::condition_variable; the uv_async hop to the V8 main loop does not exist: Manager::work calls the
// observer directly.  There is no CPU operator: every consumer needs a CUDA device (id % device count).
#pragma once
#include "../../include/tidalwave_b200.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

namespace tidalwave {

enum ErrorCode { OK, BadParameter, BadImageFormat, DontMatchSize, CudaError };

struct OpticalFlowStatus {
    ErrorCode code;
    std::string message;
    float time;
    int height;
    int width;
};

struct OpticalFlowParameter {
    double pyrScale;
    int pyrLevels;
    int winSize;
    int pyrIterations;
    int polyN;
    double polySigma;
    int flags;
};

struct Image { // cv::Mat, CV_8UC1
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;
    bool empty() const { return data.empty(); }
};

struct Plane { // cv::Mat, CV_32FC1
    int rows = 0, cols = 0;
    std::vector<float> data;
    float at(int y, int x) const { return data[(size_t)y * cols + x]; }
};

// cv::imread(path, IMREAD_GRAYSCALE): PNG and binary PGM through tw_decode_gray; anything else -> empty image
inline Image imread_gray(const std::string &path)
{
    Image img;
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return img;
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    int w = 0, h = 0;
    if (bytes.empty() || tw_decode_gray(bytes.data(), bytes.size(), NULL, 0, &w, &h) != TW_OK) return img;
    img.data.resize((size_t)w * h);
    if (tw_decode_gray(bytes.data(), bytes.size(), img.data.data(), img.data.size(), &w, &h) != TW_OK) { img.data.clear(); return img; }
    img.rows = h; img.cols = w;
    return img;
}

class OpticalFlow {
public:
    OpticalFlow() {}
    virtual ~OpticalFlow() {}

    // src/opticalflow.cpp:20-76
    OpticalFlowStatus calculate(const std::string &expectImgPath, const std::string &targetImgPath, const OpticalFlowParameter &param,
                                Plane &flowx, Plane &flowy)
    {
        OpticalFlowStatus status;
        status.code = OK; status.time = 0; status.height = 0; status.width = 0;
        this->parameter = param;
        if (expectImgPath.empty()) { status.code = BadParameter; status.message = "ExpectImagePath is empty."; return status; }
        if (targetImgPath.empty()) { status.code = BadParameter; status.message = "TargetImagePath is empty."; return status; }
        Image expectImg = imread_gray(expectImgPath);
        if (expectImg.empty()) { status.code = BadImageFormat; status.message = "Can't open " + expectImgPath; return status; }
        Image targetImg = imread_gray(targetImgPath);
        if (targetImg.empty()) { status.code = BadImageFormat; status.message = "Can't open " + targetImgPath; return status; }
        if (std::abs(expectImg.rows - targetImg.rows) > 5 || std::abs(expectImg.cols - targetImg.cols) > 5) {
            status.code = DontMatchSize; status.message = "Don't match image size"; return status;
        }
        if (expectImg.rows != targetImg.rows || expectImg.cols != targetImg.cols) {
            Image resized;
            if (!resizeTarget(targetImg, expectImg.rows, expectImg.cols, resized)) {
                status.code = CudaError; status.message = lastError(); return status;
            }
            targetImg = resized;
        }
        status.time = calculateInternal(expectImg, targetImg, flowx, flowy);
        if (status.time < 0) { status.code = failureCode(); status.message = lastError(); return status; }
        status.height = expectImg.rows;
        status.width = expectImg.cols;
        return status;
    }

protected:
    virtual float calculateInternal(Image &expectImg, Image &targetImg, Plane &flowx, Plane &flowy) = 0; // seconds, < 0 on failure
    virtual bool resizeTarget(const Image &target, int rows, int cols, Image &out) = 0;
    virtual std::string lastError() const { return std::string(); }
    virtual ErrorCode failureCode() const { return CudaError; }
    OpticalFlowParameter parameter;
};

// The B200 operator (the reference's OpticalFlowByGPU seat, src/opticalflow.cpp:97-119)
class OpticalFlowByB200 : public OpticalFlow {
public:
    explicit OpticalFlowByB200(int device) : OpticalFlow(), ctx(NULL), lastCode(TW_OK)
    {
        char err[256] = {0};
        ctx = tw_create(device, 4096, 4096, 1, err, sizeof err);
        if (!ctx) createError = err;
    }
    virtual ~OpticalFlowByB200() { tw_destroy(ctx); }
    bool ok() const { return ctx != NULL; }
    const std::string &error() const { return createError; }

protected:
    virtual float calculateInternal(Image &expectImg, Image &targetImg, Plane &flowx, Plane &flowy)
    {
        if (!ctx) { lastCode = TW_CUDA_ERROR; return -1.f; }
        tw_flow_param p = {parameter.pyrScale, parameter.pyrLevels, parameter.winSize, parameter.pyrIterations,
                           parameter.polyN, parameter.polySigma, parameter.flags};
        flowx.rows = flowy.rows = expectImg.rows; flowx.cols = flowy.cols = expectImg.cols;
        flowx.data.resize((size_t)expectImg.rows * expectImg.cols);
        flowy.data.resize(flowx.data.size());
        float seconds = 0.f;
        lastCode = tw_flow(ctx, expectImg.data.data(), targetImg.data.data(), expectImg.cols, expectImg.rows, expectImg.cols, &p,
                           flowx.data.data(), flowy.data.data(), &seconds);
        return lastCode == TW_OK ? seconds : -1.f;
    }
    virtual bool resizeTarget(const Image &target, int rows, int cols, Image &out)
    {
        if (!ctx) return false;
        out.rows = rows; out.cols = cols; out.data.resize((size_t)rows * cols);
        lastCode = tw_resize_target(ctx, target.data.data(), target.cols, target.rows, out.data.data(), cols, rows);
        return lastCode == TW_OK;
    }
    virtual std::string lastError() const { return ctx ? std::string(tw_last_error(ctx)) : createError; }
    virtual ErrorCode failureCode() const { return lastCode == TW_BAD_PARAMETER ? BadParameter : CudaError; }

private:
    tw_ctx *ctx;
    int lastCode;
    std::string createError;
};

// ---- src/message_queue.h ----
struct Request {
    std::string expect_image;
    std::string target_image;
    double threshold;
    int span;
};

struct Vector {
    int x;
    int y;
    double dx;
    double dy;
};

struct Response {
    Response() : vectors(), time(0), threshold(0), span(0), width(0), height(0) {}
    std::vector<Vector> vectors;
    std::string expect_image;
    std::string target_image;
    float time;
    double threshold;
    int span;
    std::string status;
    std::string reason;
    int width;
    int height;
};

struct Report {
    int requestCount;
    int dataCount;
    int errorCount;
};

template <typename T>
class MessageQueue {
public:
    MessageQueue() : isRunning(true) {}
    // waits for a message or the stop notice; false when stopped (src/message_queue.h:67-85)
    bool tryPop(T &buf)
    {
        std::unique_lock<std::mutex> lk(mutex);
        notifier.wait(lk, [&] { return !queue.empty() || !isRunning; });
        if (queue.empty() || !isRunning) return false;
        buf = queue.front();
        queue.pop();
        return true;
    }
    void push(const T &buf)
    {
        { std::lock_guard<std::mutex> lk(mutex); queue.push(buf); }
        notifier.notify_one();
    }
    void stop()
    {
        { std::lock_guard<std::mutex> lk(mutex); isRunning = false; }
        notifier.notify_all();
    }
    void reset() { std::lock_guard<std::mutex> lk(mutex); isRunning = true; }

private:
    std::mutex mutex;
    std::condition_variable notifier;
    bool isRunning;
    std::queue<T> queue;
    MessageQueue(const MessageQueue &);
    MessageQueue &operator=(const MessageQueue &);
};

// ---- src/observer.h ----
template <typename TValue, typename TError, typename TReport>
class Observer {
public:
    virtual ~Observer() {}
    virtual void onNext(const TValue &value) = 0;
    virtual void onError(const TError &reason) = 0;
    virtual void onCompleted(const TReport &report) = 0;
};

// ---- src/consumer.h/.cpp ----
class Consumer {
public:
    Consumer(int id, MessageQueue<Request> &reqq, MessageQueue<Response> &resq)
        : id(id), requestQueue(reqq), responseQueue(resq), isRunning(false), opticalFlow(NULL)
    {
        // consumer id <-> GPU id while id < device count (src/consumer.cpp:18-24); beyond that the devices are shared
        // round-robin -- there is no CPU operator to fall back to
        int devCount = tw_device_count();
        opticalFlow = new OpticalFlowByB200(devCount > 0 ? id % devCount : 0);
    }
    ~Consumer() { stop(); delete opticalFlow; }

    // What the worker loop of src/consumer.cpp:42-94 does, expressed over the C ABI's result types: pop, compute, answer.
    int run()
    {
        Request req;
        while (isRunning) {
            if (!requestQueue.tryPop(req)) continue;
            responseQueue.push(answer(req));
        }
        return 0;
    }

private:
    // sampling rule of src/consumer.cpp:60-77: every span-th row / column, float len, strict double compare
    static void sampleVectors(const Plane &fx, const Plane &fy, int span, double threshold, std::vector<Vector> &out)
    {
        const double limit = threshold * threshold;
        for (int y = 0; y < fx.rows; y += span)
            for (int x = 0; x < fx.cols; x += span) {
                const float dx = fx.at(y, x), dy = fy.at(y, x);
                const float len = (dx * dx) + (dy * dy);
                if (len > limit) out.push_back(Vector{x, y, dx, dy});
            }
    }
    Response answer(const Request &req)
    {
        Response res;
        Plane flowx, flowy;
        const OpticalFlowStatus st = opticalFlow->calculate(req.expect_image, req.target_image, parameter, flowx, flowy);
        if (st.code != OK) { // src/consumer.cpp:85-88
            res.status = "ERROR";
            res.reason = st.message;
            return res;
        }
        sampleVectors(flowx, flowy, req.span, req.threshold, res.vectors);
        res.status = res.vectors.empty() ? "OK" : "SUSPICIOUS";
        res.expect_image = req.expect_image; res.target_image = req.target_image;
        res.span = req.span; res.threshold = req.threshold;
        res.time = st.time; res.height = st.height; res.width = st.width;
        return res;
    }

public:
    void start(const OpticalFlowParameter &param)
    {
        parameter = param;
        isRunning = true;
        thread = std::thread([this] { run(); });
    }
    int stop()
    {
        isRunning = false;
        if (thread.joinable()) thread.join();
        return 0;
    }

private:
    int id;
    MessageQueue<Request> &requestQueue;
    MessageQueue<Response> &responseQueue;
    std::atomic<bool> isRunning;
    OpticalFlowParameter parameter;
    OpticalFlowByB200 *opticalFlow;
    std::thread thread;
};

// ---- src/manager.h/.cpp ----
struct Parameter {
    double threshold;
    int span;
    int numThreads;
    OpticalFlowParameter optParam;
};

class Manager {
public:
    explicit Manager(Observer<Response, std::string, Report> *emitter) : isRunning(true), emitter(emitter), answered(0)
    {
        report.requestCount = report.dataCount = report.errorCount = 0;
    }
    virtual ~Manager()
    {
        if (worker.joinable()) { stop(); worker.join(); }
        for (size_t i = 0; i < consumers.size(); i++) delete consumers[i];
    }
    int start(const Parameter &p) // src/manager.cpp:40-61
    {
        param = p;
        requestQueue.reset();
        responseQueue.reset();
        isRunning = true;
        worker = std::thread([this] { work(); finish(); });
        for (int i = 0; i < param.numThreads; i++) {
            Consumer *cons = new Consumer(i, requestQueue, responseQueue);
            cons->start(param.optParam);
            consumers.push_back(cons);
        }
        return 0;
    }
    void stop() // src/manager.cpp:63-66
    {
        isRunning = false;
        responseQueue.stop();
    }
    int request(const std::string &expect_image, const std::string &target_image) // src/manager.cpp:68-78
    {
        Request req;
        req.expect_image = expect_image;
        req.target_image = target_image;
        req.span = param.span;
        req.threshold = param.threshold;
        requestQueue.push(req);
        std::lock_guard<std::mutex> lk(mu);
        report.requestCount++;
        return 0;
    }
    void work() // src/manager.cpp:80-98 + notify :102-125 (no V8 main-loop hop)
    {
        while (isRunning) {
            Response res;
            if (responseQueue.tryPop(res)) {
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (res.status == "ERROR") report.errorCount++; else report.dataCount++;
                    answered++;
                }
                if (res.status == "ERROR") emitter->onError(res.reason); else emitter->onNext(res);
                cv.notify_all();
            }
        }
        requestQueue.stop();
        for (size_t i = 0; i < consumers.size(); i++) consumers[i]->stop();
    }
    void finish() { emitter->onCompleted(report); } // src/manager.cpp:128-133
    // helper for hosts without an event loop: index.js:64-68 disposes once every request was answered
    void waitAllAnswered()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return answered >= report.requestCount; });
    }
    void join() { if (worker.joinable()) worker.join(); }

    std::atomic<bool> isRunning;

private:
    MessageQueue<Request> requestQueue;
    MessageQueue<Response> responseQueue;
    std::vector<Consumer *> consumers;
    Parameter param;
    Report report;
    Observer<Response, std::string, Report> *emitter;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    int answered;
};

} // namespace tidalwave
