// ros_bus_check.cpp -- the roscpp stand-in's bus mode (oracle/ref_shim/ros/ros.h), which tests/native/boloop relies on: two
// nodes of one process, one thread each; publish() reaches the process' own subscribers, spinOnce() delivers to the callbacks
// the CALLING thread registered, spin() serves until ros::shutdown().  A ping-pong of N messages must come back complete, in
// order, each callback on its own node's thread.
#include <cstdio>
#include <thread>

#include <ros/ros.h>
#include <std_msgs/Float64.h>

namespace {
struct Echo {                                    // node A: answers every "ping" with "pong" = 2 x ping
    ros::Publisher pub;
    std::thread::id seen_on;
    int n = 0;
    void OnPing(const std_msgs::Float64::ConstPtr &m)
    {
        seen_on = std::this_thread::get_id();
        std_msgs::Float64 r;
        r.data = 2.0 * m->data;
        ++n;
        pub.publish(r);
    }
};
struct Caller {                                  // node B: publishes "ping", spins until its "pong" arrives
    double answer = -1.0;
    std::thread::id seen_on;
    void OnPong(const std_msgs::Float64::ConstPtr &m) { answer = m->data; seen_on = std::this_thread::get_id(); }
};
}  // namespace

int main()
{
    const int N = 2000;
    ros::ShimMaster::get().set_bus(true);
    Echo echo;
    std::thread::id a_id;
    std::thread a([&] {
        a_id = std::this_thread::get_id();
        ros::NodeHandle n;
        echo.pub = n.advertise<std_msgs::Float64>("pong", 10);
        ros::Subscriber s = n.subscribe<std_msgs::Float64>("ping", 10, &Echo::OnPing, &echo);
        (void)s;
        ros::spin();
    });
    while (ros::ShimMaster::get().subscribers("ping") == 0) std::this_thread::yield();
    ros::NodeHandle n;
    Caller caller;
    ros::Publisher ping = n.advertise<std_msgs::Float64>("ping", 10);
    ros::Subscriber s = n.subscribe<std_msgs::Float64>("pong", 10, &Caller::OnPong, &caller);
    (void)s;
    int bad = 0;
    for (int i = 0; i < N; ++i) {
        caller.answer = -1.0;
        std_msgs::Float64 m;
        m.data = (double)i;
        ping.publish(m);
        while (caller.answer == -1.0 && ros::ok()) ros::spinOnce();     // robot1d_bayesopt.cpp:107-109
        if (caller.answer != 2.0 * i) ++bad;
    }
    ros::shutdown();
    a.join();
    const bool threads_ok = echo.seen_on == a_id && caller.seen_on == std::this_thread::get_id();
    std::printf("%s: %d round trips, %d wrong, echo served %d, callbacks on their own threads: %s\n",
                (bad == 0 && echo.n == N && threads_ok) ? "ok ros_bus" : "FAILED ros_bus", N, bad, echo.n, threads_ok ? "yes" : "no");
    return (bad == 0 && echo.n == N && threads_ok) ? 0 : 1;
}
